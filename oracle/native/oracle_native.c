/*
 * oracle_native.c -- CPU restatement (plain C, fp32) of the reference's first-party native code.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Build: oracle/build_oracle.py (gcc -O2
 * -ffp-contract=off, no fast-math: every product and sum rounds separately, as the reference's
 * host code compiled for baseline x86-64 does).
 *
 * Restated from (paths relative to the reference repository):
 *   rotated rectangle overlap / BEV IoU   models/bbox_post_process/src/iou3d_nms_kernel.cu:34-234
 *                                          (host twin models/bbox_post_process/src/iou3d_cpu.cpp:38-228)
 *   64x64 suppression mask                 iou3d_nms_kernel.cu:267-311
 *   greedy sweep                           models/bbox_post_process/src/iou3d_nms.cpp:118-132
 *   voxel-in-box membership                models/utils/src/Array_Index.cpp:14-79
 * Pinned against the compiled reference by tests/test_oracle_native.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y; } pt;

static float cross_o(pt p1, pt p2, pt p0) {            /* iou3d_nms_kernel.cu:38-40 */
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}
static float cross_v(pt a, pt b) { return a.x * b.y - a.y * b.x; }   /* :34-36 */
static float fmin2(float a, float b) { return a < b ? a : b; }
static float fmax2(float a, float b) { return a > b ? a : b; }

static int rect_cross(pt p1, pt p2, pt q1, pt q2) {    /* :42-48 */
    return fmin2(p1.x, p2.x) <= fmax2(q1.x, q2.x) && fmin2(q1.x, q2.x) <= fmax2(p1.x, p2.x) &&
           fmin2(p1.y, p2.y) <= fmax2(q1.y, q2.y) && fmin2(q1.y, q2.y) <= fmax2(p1.y, p2.y);
}

static int in_box2d(const float* box, pt p) {          /* :50-60, MARGIN 1e-2 */
    const float margin = 1e-2f;
    float cx = box[0], cy = box[1];
    float ac = cosf(-box[6]), as = sinf(-box[6]);
    float rx = (p.x - cx) * ac + (p.y - cy) * (-as);
    float ry = (p.x - cx) * as + (p.y - cy) * ac;
    return fabsf(rx) < box[3] / 2 + margin && fabsf(ry) < box[4] / 2 + margin;
}

static int seg_x(pt p1, pt p0, pt q1, pt q0, pt* ans) {   /* :62-93 */
    if (!rect_cross(p0, p1, q0, q1)) return 0;
    float s1 = cross_o(q0, p1, p0), s2 = cross_o(p1, q1, p0);
    float s3 = cross_o(p0, q1, q0), s4 = cross_o(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
    float s5 = cross_o(q1, p1, p0);
    if (fabsf(s5 - s1) > 1e-8f) {
        ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        float D = a0 * b1 - a1 * b0;
        ans->x = (b0 * c1 - b1 * c0) / D;
        ans->y = (a1 * c0 - a0 * c1) / D;
    }
    return 1;
}

static void corners(const float* b, pt c[5]) {          /* :108-150 */
    float hx = b[3] / 2, hy = b[4] / 2;
    float x1 = b[0] - hx, y1 = b[1] - hy, x2 = b[0] + hx, y2 = b[1] + hy;
    float ca = cosf(b[6]), sa = sinf(b[6]);
    float px[4] = {x1, x2, x2, x1}, py[4] = {y1, y1, y2, y2};
    for (int k = 0; k < 4; ++k) {                      /* rotate_around_center :95-99 */
        c[k].x = (px[k] - b[0]) * ca + (py[k] - b[1]) * (-sa) + b[0];
        c[k].y = (px[k] - b[0]) * sa + (py[k] - b[1]) * ca + b[1];
    }
    c[4] = c[0];
}

float oracle_box_overlap(const float* A, const float* B) {   /* :104-225 */
    pt ca[5], cb[5], pts[16], ctr = {0.f, 0.f};
    corners(A, ca); corners(B, cb);
    int cnt = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            pt x;
            if (seg_x(ca[i + 1], ca[i], cb[j + 1], cb[j], &x)) { ctr.x += x.x; ctr.y += x.y; pts[cnt++] = x; }
        }
    for (int k = 0; k < 4; ++k) {
        if (in_box2d(A, cb[k])) { ctr.x += cb[k].x; ctr.y += cb[k].y; pts[cnt++] = cb[k]; }
        if (in_box2d(B, ca[k])) { ctr.x += ca[k].x; ctr.y += ca[k].y; pts[cnt++] = ca[k]; }
    }
    ctr.x /= cnt; ctr.y /= cnt;
    for (int j = 0; j < cnt - 1; ++j)                   /* bubble sort by angle :196-207 */
        for (int i = 0; i < cnt - j - 1; ++i)
            if (atan2f(pts[i].y - ctr.y, pts[i].x - ctr.x) > atan2f(pts[i + 1].y - ctr.y, pts[i + 1].x - ctr.x)) {
                pt t = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = t;
            }
    float area = 0.f;
    for (int k = 0; k < cnt - 1; ++k) {
        pt u = {pts[k].x - pts[0].x, pts[k].y - pts[0].y}, v = {pts[k + 1].x - pts[0].x, pts[k + 1].y - pts[0].y};
        area += cross_v(u, v);
    }
    return fabsf(area) / 2.0f;
}

float oracle_iou_bev(const float* A, const float* B) {       /* :227-234 */
    float sa = A[3] * A[4], sb = B[3] * B[4];
    float so = oracle_box_overlap(A, B);
    return so / fmaxf(sa + sb - so, 1e-8f);
}

void oracle_iou_matrix(const float* a, int na, const float* b, int nb, float* out) {
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) out[(size_t)i * nb + j] = oracle_iou_bev(a + i * 7, b + j * 7);
}

/* nms_gpu: boxes sorted by descending score.  keep[] gets ALL kept indices (ascending); returns count. */
int oracle_nms(const float* boxes, int n, float thresh, int64_t* keep) {
    const int cb = (n + 63) / 64;
    uint64_t* mask = (uint64_t*)calloc((size_t)n * cb + 1, sizeof(uint64_t));
    /* every 64-bit mask word is independent (the reference computes them in parallel CUDA blocks): rows in parallel */
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = 0; i < n; ++i)                           /* nms_kernel :267-311 */
        for (int c = i / 64; c < cb; ++c) {
            uint64_t t = 0;
            int start = (c == i / 64) ? (i % 64) + 1 : 0;
            int ncol = n - c * 64 < 64 ? n - c * 64 : 64;
            for (int j = start; j < ncol; ++j)
                if (oracle_iou_bev(boxes + (size_t)i * 7, boxes + (size_t)(c * 64 + j) * 7) > thresh) t |= 1ull << j;
            mask[(size_t)i * cb + c] = t;
        }
    uint64_t* remv = (uint64_t*)calloc(cb + 1, sizeof(uint64_t));
    int num = 0;
    for (int i = 0; i < n; ++i) {                         /* iou3d_nms.cpp:118-132 */
        int nb = i / 64, ib = i % 64;
        if (!(remv[nb] & (1ull << ib))) {
            keep[num++] = i;
            const uint64_t* p = mask + (size_t)i * cb;
            for (int j = nb; j < cb; ++j) remv[j] |= p[j];
        }
    }
    free(mask); free(remv);
    return num;
}

/* Array_Index.find_features_by_bbox_with_yaw: vox [n,3] (x,y,z) int32, boxes [nb,8] float, out [n,ncls] int32 */
void oracle_find_features_by_bbox_with_yaw(const int32_t* vox, int n, const float* boxes, int nb, int32_t* out, int ncls) {
    for (int i = 0; i < nb; ++i) {
        const float* b = boxes + (size_t)i * 8;
        float center[3] = {b[0], b[1], b[2]}, extend[3] = {b[3], b[4], b[5]};
        float theta = b[6];
        float cos_theta = cosf(theta), sin_theta = sinf(theta);
        int first_flag = 0, first_point[3] = {0, 0, 0};
        for (int j = 0; j < n; ++j) {
            const int32_t* r = vox + (size_t)j * 3;
            if (first_flag == 1 &&
                (r[0] > (first_point[0] + extend[0]) || r[0] < (first_point[0] - extend[0]) ||
                 r[1] > (first_point[1] + extend[1]) || r[1] < (first_point[1] - extend[1]) ||
                 r[2] > (first_point[2] + extend[2]) || r[2] < (first_point[2] - extend[2])))
                continue;
            float centered[3] = {r[0] - center[0], r[1] - center[1], r[2] - center[2]};
            float rx = centered[0] * cos_theta + centered[1] * sin_theta;
            float ry = -centered[0] * sin_theta + centered[1] * cos_theta;
            if (rx <= extend[0] / 2 && rx >= -extend[0] / 2 && ry <= extend[1] / 2 && ry >= -extend[1] / 2 &&
                centered[2] <= extend[2] / 2 && centered[2] >= -extend[2] / 2) {
                int label = (int)b[7];
                if (label > 0 && label <= ncls) out[(size_t)j * ncls + label - 1] = 1;
                if (!first_flag) { first_flag = 1; first_point[0] = r[0]; first_point[1] = r[1]; first_point[2] = r[2]; }
            }
        }
    }
}

/* Array_Index.find_point_in_instance_bbox_with_yaw (models/utils/src/Array_Index.cpp:83-149), boxes visited in ascending
 * order (the reference's OpenMP loop over boxes races when two boxes of one class contain the same point; serial order =
 * the later box wins).  pts [n,stride] float (x,y,z first), boxes [nb,8] (cx,cy,cz,dx,dy,dz,yaw,label), out [n,ncls] int32
 * (caller zeroes it): out[j,label-1] = box index + 1. */
void oracle_find_point_in_instance_bbox_with_yaw(const float* pts, int n, int stride, const float* boxes, int nb,
                                                 int32_t* out, int ncls, float out_ground) {
    for (int i = 0; i < nb; ++i) {
        const float* b = boxes + (size_t)i * 8;
        float center[3] = {b[0], b[1], b[2] + out_ground}, extend[3] = {b[3], b[4], b[5]};
        float theta = b[6];
        float cos_theta = cosf(theta), sin_theta = sinf(theta);
        int first_flag = 0;
        float first_point[3] = {0.f, 0.f, 0.f};
        for (int j = 0; j < n; ++j) {
            const float* r = pts + (size_t)j * stride;
            if (first_flag == 1 &&
                (r[0] > (first_point[0] + extend[0]) || r[0] < (first_point[0] - extend[0]) ||
                 r[1] > (first_point[1] + extend[1]) || r[1] < (first_point[1] - extend[1]) ||
                 r[2] > (first_point[2] + extend[2]) || r[2] < (first_point[2] - extend[2])))
                continue;
            float centered[3] = {r[0] - center[0], r[1] - center[1], r[2] - center[2]};
            float rx = centered[0] * cos_theta + centered[1] * sin_theta;
            float ry = -centered[0] * sin_theta + centered[1] * cos_theta;
            if (rx <= extend[0] / 2 && rx >= -extend[0] / 2 && ry <= extend[1] / 2 && ry >= -extend[1] / 2 &&
                centered[2] <= extend[2] / 2 && centered[2] >= -extend[2] / 2) {
                int label = (int)b[7];
                if (label > 0 && label <= ncls) out[(size_t)j * ncls + label - 1] = i + 1;
                if (!first_flag) { first_flag = 1; first_point[0] = r[0]; first_point[1] = r[1]; first_point[2] = r[2]; }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * Kernel-map neighbour table (MinkowskiEngine CPU coordinate map / spconv CPU indice pairs, restated: both libraries keep
 * the input coordinates in a hash map and, for every kernel offset, probe `query + offset` for every query row; ME runs the
 * offsets in parallel with OpenMP).  Same result as oracle/me.py::kernel_map / oracle/sp.py::subm_maps, which remain the
 * plain-numpy statement (sort + binary search) and the checker of this function (tests/test_oracle_ops.py).
 *
 *   in_coords  [n_in, ncol]  int64 rows (batch, c0, c1, c2[, c3]); duplicates: the FIRST row wins (as np.searchsorted on a
 *                            stable sort does)
 *   q_coords   [n_q, ncol]   query rows
 *   offs       [K, D]        per-offset deltas, added to columns 1 .. D
 *   nbr        [K, n_q]      out: row of `q + offs[k]` in in_coords, or -1 (also -1 when the shifted coordinate leaves the
 *                            packable range |c0..c2| < 2^15, |c3| < 128 -- oracle/me.py::kernel_map's `ok` mask)
 * Key packing = oracle/me.py::pack_keys. */

#ifdef _OPENMP
#include <omp.h>
#endif
/* threads of the parallel loops below (torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU baseline sets its own count) */
void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static inline uint64_t okm_mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
static inline int64_t okm_pack(int64_t b, int64_t c0, int64_t c1, int64_t c2, int64_t c3) {
    const int64_t B = 1 << 15;
    int64_t key = b;
    key = key * (2 * B) + (c0 + B);
    key = key * (2 * B) + (c1 + B);
    key = key * (2 * B) + (c2 + B);
    return key * 256 + (c3 + 128);
}

int oracle_neighbor_table(const int64_t* in_coords, int64_t n_in, const int64_t* q_coords, int64_t n_q, int ncol,
                          const int64_t* offs, int K, int D, int32_t* nbr) {
    if (ncol < 4 || ncol > 5 || D < 1 || D > ncol - 1 || n_in < 0 || n_q < 0 || K < 0 || n_in > 0x7fffffffLL) return -1;
    uint64_t cap = 16;
    while (cap < 2 * (uint64_t)n_in) cap <<= 1;
    const uint64_t mask = cap - 1;
    int64_t* keys = (int64_t*)malloc(cap * sizeof(int64_t));
    int32_t* rows = (int32_t*)malloc(cap * sizeof(int32_t));
    if (!keys || !rows) { free(keys); free(rows); return -2; }
    for (uint64_t i = 0; i < cap; ++i) keys[i] = -1;                       /* packed keys of valid rows are >= 0 */
    for (int64_t i = 0; i < n_in; ++i) {                                    /* sequential insert: the first row of a key wins */
        const int64_t* c = in_coords + i * ncol;
        const int64_t key = okm_pack(c[0], c[1], c[2], c[3], ncol > 4 ? c[4] : 0);
        uint64_t s = okm_mix((uint64_t)key) & mask;
        while (keys[s] != -1 && keys[s] != key) s = (s + 1) & mask;
        if (keys[s] == -1) { keys[s] = key; rows[s] = (int32_t)i; }
    }
    const int64_t B = 1 << 15;
#pragma omp parallel for schedule(dynamic, 1)
    for (int k = 0; k < K; ++k) {
        int64_t d[4] = {0, 0, 0, 0};
        for (int j = 0; j < D; ++j) d[j] = offs[(int64_t)k * D + j];
        int32_t* out = nbr + (int64_t)k * n_q;
        for (int64_t i = 0; i < n_q; ++i) {
            const int64_t* c = q_coords + i * ncol;
            const int64_t c0 = c[1] + d[0], c1 = c[2] + d[1], c2 = c[3] + d[2], c3 = (ncol > 4 ? c[4] : 0) + d[3];
            int32_t r = -1;
            if (c0 > -B && c0 < B && c1 > -B && c1 < B && c2 > -B && c2 < B && (ncol <= 4 || (c3 > -128 && c3 < 128))) {
                const int64_t key = okm_pack(c[0], c0, c1, c2, c3);
                uint64_t s = okm_mix((uint64_t)key) & mask;
                while (keys[s] != -1) {
                    if (keys[s] == key) { r = rows[s]; break; }
                    s = (s + 1) & mask;
                }
            }
            out[i] = r;
        }
    }
    free(keys);
    free(rows);
    return 0;
}
