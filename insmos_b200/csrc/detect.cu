// Detection head pieces of the hot path:
//   * CenterHead decode fused with sigmoid / class max            (center_head.py:251-276, post_process.py:146-192)
//   * rotated BEV NMS: 64x64 IoU bit-mask + DEVICE-side greedy sweep  (iou3d_nms_kernel.cu:104-311, iou3d_nms.cpp:90-136)
//     -- the reference copies the mask to the host and sweeps there (a sync per forward); here the
//        sweep is a second kernel, so the forward has no host round trip for NMS.
//   * voxel-in-rotated-box membership, reproducing Array_Index.cpp:14-79 including its first-hit
//     pruning, on device (the reference does 4 D2H + OpenMP + H2D round trips per forward).
// The geometry follows the reference's arithmetic step by step in fp32 (this file is compiled with
// -fmad=false so products and sums round separately like the host twin iou3d_cpu.cpp / the
// oracle); decisions (suppress / keep / inside) must agree with the oracle, not just areas.
#include "common.cuh"
#include <math.h>

// ------------------------------------------------------------------------------------------------
#define BOXV(c) box[(int64_t)(c) * box_cs + (int64_t)i * box_ps]
__global__ void k_center_decode(const float* __restrict__ cls, int64_t cls_cs, int64_t cls_ps,
                                const float* __restrict__ box, int64_t box_cs, int64_t box_ps, int ncls, int H, int W,
                                float osf, float vx, float vy, float x_min, float y_min,
                                float* __restrict__ boxes, float* __restrict__ scores, int32_t* __restrict__ labels) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int HW = H * W;
    if (i >= HW) return;
    const int y = i / W, x = i - y * W;
    // xs = (x + reg0) * OUT_SIZE_FACTOR * VOXEL_SIZE[0] + range_min   (center_head.py:264-268), fp32, left to right
    float xs = __fadd_rn((float)x, BOXV(0));
    float ys = __fadd_rn((float)y, BOXV(1));
    xs = __fadd_rn(__fmul_rn(__fmul_rn(xs, osf), vx), x_min);
    ys = __fadd_rn(__fmul_rn(__fmul_rn(ys, osf), vy), y_min);
    float* o = boxes + (int64_t)i * 7;
    o[0] = xs; o[1] = ys; o[2] = BOXV(2);
    o[3] = expf(BOXV(3)); o[4] = expf(BOXV(4)); o[5] = expf(BOXV(5));
    o[6] = atan2f(BOXV(6), BOXV(7));
    float best = -1.0f; int bi = 0;
    for (int c = 0; c < ncls; ++c) {
        const float l = cls[(int64_t)c * cls_cs + (int64_t)i * cls_ps];
        const float s = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-l)));      // torch.sigmoid
        if (s > best) { best = s; bi = c; }                              // first maximum wins (torch.max)
    }
    scores[i] = best; labels[i] = bi + 1;
}
extern "C" int insmos_center_decode(const float* cls, int64_t cls_cs, int64_t cls_ps, const float* box, int64_t box_cs,
                                    int64_t box_ps, int32_t ncls, int32_t H, int32_t W,
                                    float out_size_factor, float vx, float vy, float x_min, float y_min,
                                    float* boxes, float* scores, int32_t* labels, void* stream) {
    if (!cls || !box || !boxes || !scores || !labels || ncls <= 0 || H <= 0 || W <= 0) return INSMOS_ERR_INVALID_ARG;
    k_center_decode<<<(H * W + 255) / 256, 256, 0, (cudaStream_t)stream>>>(cls, cls_cs, cls_ps, box, box_cs, box_ps, ncls, H, W,
                                                                            out_size_factor, vx, vy,
                                                                            x_min, y_min, boxes, scores, labels);
    INSMOS_CHECK_LAUNCH("k_center_decode");
    return INSMOS_OK;
}

// ------------------------------------------------------------------------------------------------
// rotated rectangle overlap, same construction as iou3d_nms_kernel.cu:104-225:
// edge-edge intersections (strict straddle test), corners of one box inside the other with a
// 1e-2 margin, centroid, bubble sort by atan2 about the centroid, shoelace area.
struct P2 { float x, y; };
__device__ __forceinline__ float cross3(const P2& p1, const P2& p2, const P2& p0) {
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}
__device__ __forceinline__ float cross2(const P2& a, const P2& b) { return a.x * b.y - a.y * b.x; }

__device__ __forceinline__ bool seg_intersect(const P2& p1, const P2& p0, const P2& q1, const P2& q0, P2& ans) {
    // bounding-box rejection
    if (!(fminf(p0.x, p1.x) <= fmaxf(q0.x, q1.x) && fminf(q0.x, q1.x) <= fmaxf(p0.x, p1.x) &&
          fminf(p0.y, p1.y) <= fmaxf(q0.y, q1.y) && fminf(q0.y, q1.y) <= fmaxf(p0.y, p1.y)))
        return false;
    const float s1 = cross3(q0, p1, p0), s2 = cross3(p1, q1, p0);
    const float s3 = cross3(p0, q1, q0), s4 = cross3(q1, p1, q0);
    if (!(s1 * s2 > 0.0f && s3 * s4 > 0.0f)) return false;
    const float s5 = cross3(q1, p1, p0);
    if (fabsf(s5 - s1) > 1e-8f) {
        ans.x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans.y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        const float D = a0 * b1 - a1 * b0;
        ans.x = (b0 * c1 - b1 * c0) / D;
        ans.y = (a1 * c0 - a0 * c1) / D;
    }
    return true;
}
__device__ __forceinline__ bool corner_in_box(const float* box, const P2& p) {
    const float ac = cosf(-box[6]), as = sinf(-box[6]);
    const float rx = (p.x - box[0]) * ac + (p.y - box[1]) * (-as);
    const float ry = (p.x - box[0]) * as + (p.y - box[1]) * ac;
    return fabsf(rx) < box[3] / 2 + 1e-2f && fabsf(ry) < box[4] / 2 + 1e-2f;
}
__device__ __forceinline__ void box_corners(const float* b, P2 (&c)[5]) {
    const float hx = b[3] / 2, hy = b[4] / 2;
    const float x1 = b[0] - hx, y1 = b[1] - hy, x2 = b[0] + hx, y2 = b[1] + hy;
    const float ca = cosf(b[6]), sa = sinf(b[6]);
    const float px[4] = {x1, x2, x2, x1}, py[4] = {y1, y1, y2, y2};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c[k].x = (px[k] - b[0]) * ca + (py[k] - b[1]) * (-sa) + b[0];
        c[k].y = (px[k] - b[0]) * sa + (py[k] - b[1]) * ca + b[1];
    }
    c[4] = c[0];
}
__device__ float rot_overlap(const float* A, const float* B) {
    P2 ca[5], cb[5];
    box_corners(A, ca); box_corners(B, cb);
    P2 pts[16]; P2 ctr; ctr.x = 0.f; ctr.y = 0.f;
    int cnt = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            P2 x;
            if (seg_intersect(ca[i + 1], ca[i], cb[j + 1], cb[j], x)) { pts[cnt++] = x; ctr.x += x.x; ctr.y += x.y; }
        }
    for (int k = 0; k < 4; ++k) {
        if (corner_in_box(A, cb[k])) { ctr.x += cb[k].x; ctr.y += cb[k].y; pts[cnt++] = cb[k]; }
        if (corner_in_box(B, ca[k])) { ctr.x += ca[k].x; ctr.y += ca[k].y; pts[cnt++] = ca[k]; }
    }
    ctr.x /= cnt; ctr.y /= cnt;                                   // cnt == 0 -> NaN, unused below
    for (int j = 0; j < cnt - 1; ++j)
        for (int i = 0; i < cnt - j - 1; ++i)
            if (atan2f(pts[i].y - ctr.y, pts[i].x - ctr.x) > atan2f(pts[i + 1].y - ctr.y, pts[i + 1].x - ctr.x)) {
                const P2 tmp = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = tmp;
            }
    float area = 0.f;
    for (int k = 0; k < cnt - 1; ++k) {
        P2 u, v;
        u.x = pts[k].x - pts[0].x; u.y = pts[k].y - pts[0].y;
        v.x = pts[k + 1].x - pts[0].x; v.y = pts[k + 1].y - pts[0].y;
        area += cross2(u, v);
    }
    return fabsf(area) / 2.0f;
}
__device__ __forceinline__ float rot_iou_bev(const float* A, const float* B) {
    const float sa = A[3] * A[4], sb = B[3] * B[4];
    const float so = rot_overlap(A, B);
    return so / fmaxf(sa + sb - so, 1e-8f);
}

// iou3d_nms_utils.boxes_iou3d_gpu (iou3d_nms_utils.py:27-61) in one kernel: rotated BEV overlap (iou3d_nms_kernel.cu:236-249)
// x height overlap / (vol_a + vol_b - overlap), the torch expressions in their fp32 order.  One thread per (a, b) pair.
__global__ void k_boxes_iou3d(const float* __restrict__ a, int na, const float* __restrict__ b, int nb, float* __restrict__ iou) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= na || j >= nb) return;
    const float* A = a + (size_t)i * 7;
    const float* B = b + (size_t)j * 7;
    const float bev = rot_overlap(A, B);
    const float a_max = A[2] + A[5] / 2.0f, a_min = A[2] - A[5] / 2.0f;
    const float b_max = B[2] + B[5] / 2.0f, b_min = B[2] - B[5] / 2.0f;
    const float h = fmaxf(fminf(a_max, b_max) - fmaxf(a_min, b_min), 0.0f);
    const float o3d = bev * h;
    const float va = A[3] * A[4] * A[5], vb = B[3] * B[4] * B[5];
    iou[(size_t)i * nb + j] = o3d / fmaxf(va + vb - o3d, 1e-6f);
}

extern "C" int insmos_boxes_iou3d(const float* boxes_a, int32_t na, const float* boxes_b, int32_t nb, float* iou, void* stream) {
    if (na < 0 || nb < 0 || (na > 0 && nb > 0 && (!boxes_a || !boxes_b || !iou))) return INSMOS_ERR_INVALID_ARG;
    if (na == 0 || nb == 0) return INSMOS_OK;
    const dim3 threads(16, 16), blocks((nb + 15) / 16, (na + 15) / 16);
    k_boxes_iou3d<<<blocks, threads, 0, (cudaStream_t)stream>>>(boxes_a, na, boxes_b, nb, iou);
    INSMOS_CHECK_LAUNCH("k_boxes_iou3d");
    return INSMOS_OK;
}

// mask[i*cb + c] bit j = IoU(box i, box c*64+j) > thresh, only for j > i (iou3d_nms_kernel.cu:267-311)
__global__ void __launch_bounds__(64)
k_nms_mask(const float* __restrict__ boxes, int n, float thresh, unsigned long long* __restrict__ mask) {
    const int rb = blockIdx.y, cb = blockIdx.x;
    const int col_blocks = (n + 63) / 64;
    __shared__ float sb[64 * 7];
    const int ncol = min(64, n - cb * 64), nrow = min(64, n - rb * 64);
    if (threadIdx.x < ncol) {
        const float* s = boxes + (int64_t)(cb * 64 + threadIdx.x) * 7;
#pragma unroll
        for (int k = 0; k < 7; ++k) sb[threadIdx.x * 7 + k] = s[k];
    }
    __syncthreads();
    if (threadIdx.x >= nrow) return;
    const int i = rb * 64 + threadIdx.x;
    unsigned long long bits = 0ull;
    if (cb >= rb) {                                               // lower triangle is never read by the sweep
        float a[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) a[k] = boxes[(int64_t)i * 7 + k];
        const int start = (rb == cb) ? threadIdx.x + 1 : 0;
        // Exact early-out: if the centres are farther apart than the two half-diagonals plus the 1e-2 corner
        // margin (with slack), no edge pair can intersect and no corner can pass check_in_box2d, so the
        // reference's construction yields cnt == 0 -> overlap exactly 0 -> never > thresh (thresh >= 0).
        const float ra = 0.5f * sqrtf(a[3] * a[3] + a[4] * a[4]);
        for (int j = start; j < ncol; ++j) {
            const float* b = sb + j * 7;
            if (thresh >= 0.0f) {
                const float dx = a[0] - b[0], dy = a[1] - b[1];
                const float reach = (ra + 0.5f * sqrtf(b[3] * b[3] + b[4] * b[4])) * 1.001f + 0.05f;
                if (dx * dx + dy * dy > reach * reach) continue;
            }
            if (rot_iou_bev(a, b) > thresh) bits |= 1ull << j;
        }
    }
    mask[(int64_t)i * col_blocks + cb] = bits;
}

// ---- pair-list variant of the mask build --------------------------------------------------------
// ncu of k_nms_mask (profiles/r01_fwd_v7_summary.txt): 0.30 ms at 18 % warps active.  A thread owns a row and walks 64
// columns; ~0.5 % of the (i, j) pairs pass the centre-distance test, so the ~4 k-instruction rotated-overlap code runs
// for ONE lane of a warp at a time (40 k executions per 4096 boxes).  Here the distance test only APPENDS the surviving
// pair to a list, and a second kernel evaluates one pair per thread (32 per warp) and sets the mask bit with atomicOr:
// the same bits, ~30x fewer executions of the expensive path.  If the list overflows (degenerate input: everything
// overlaps everything) the dense kernel runs instead -- decided on the device, no host read.
__global__ void __launch_bounds__(64)
k_nms_pairs(const float* __restrict__ boxes, int n, uint2* __restrict__ pairs, unsigned int cap, unsigned int* __restrict__ count) {
    const int rb = blockIdx.y, cb = blockIdx.x;
    if (cb < rb) return;                                          // lower triangle is never read by the sweep
    __shared__ float sb[64 * 7];
    const int ncol = min(64, n - cb * 64), nrow = min(64, n - rb * 64);
    if (threadIdx.x < ncol) {
        const float* s = boxes + (int64_t)(cb * 64 + threadIdx.x) * 7;
#pragma unroll
        for (int k = 0; k < 7; ++k) sb[threadIdx.x * 7 + k] = s[k];
    }
    __syncthreads();
    if (threadIdx.x >= nrow) return;
    const int i = rb * 64 + threadIdx.x;
    const float ax = boxes[(int64_t)i * 7], ay = boxes[(int64_t)i * 7 + 1], aw = boxes[(int64_t)i * 7 + 3], al = boxes[(int64_t)i * 7 + 4];
    const float ra = 0.5f * sqrtf(aw * aw + al * al);
    const int start = (rb == cb) ? threadIdx.x + 1 : 0;
    for (int j = start; j < ncol; ++j) {
        const float* b = sb + j * 7;
        const float dx = ax - b[0], dy = ay - b[1];
        const float reach = (ra + 0.5f * sqrtf(b[3] * b[3] + b[4] * b[4])) * 1.001f + 0.05f;     // same exact early-out as k_nms_mask
        if (dx * dx + dy * dy > reach * reach) continue;
        const unsigned int pos = atomicAdd(count, 1u);
        if (pos < cap) pairs[pos] = make_uint2((unsigned)i, (unsigned)(cb * 64 + j));
    }
}
__global__ void __launch_bounds__(128)
k_nms_pair_iou(const float* __restrict__ boxes, int n, float thresh, const uint2* __restrict__ pairs, unsigned int cap,
               const unsigned int* __restrict__ count, unsigned long long* __restrict__ mask) {
    const unsigned int total = *count;
    if (total > cap) return;                                      // overflow: the dense kernel takes over
    const int col_blocks = (n + 63) / 64;
    for (unsigned int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
        const uint2 ij = pairs[p];
        float a[7], b[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) { a[k] = boxes[(int64_t)ij.x * 7 + k]; b[k] = boxes[(int64_t)ij.y * 7 + k]; }
        if (rot_iou_bev(a, b) > thresh)
            atomicOr(mask + (int64_t)ij.x * col_blocks + (ij.y >> 6), 1ull << (ij.y & 63u));
    }
}
// dense fallback, skipped unless the pair list overflowed
__global__ void __launch_bounds__(64)
k_nms_mask_if_overflow(const float* __restrict__ boxes, int n, float thresh, unsigned long long* __restrict__ mask,
                       const unsigned int* __restrict__ count, unsigned int cap);

// greedy sweep on device: 64 boxes at a time.  Thread 0 resolves a block from the diagonal words held in shared memory
// (branch-free body: the 64 diagonal loads do not depend on the decisions), then 256 threads OR the kept rows into the
// removal words of the later blocks: thread = (column word, quarter of the kept rows), all loads of a quarter issued before
// any is consumed.  Stops as soon as max_keep boxes are kept (only keep[:NMS_POST_MAXSIZE] is ever used: post_process.py:17).
#define SW_THREADS 256
__global__ void __launch_bounds__(SW_THREADS)
k_nms_sweep(const unsigned long long* __restrict__ mask, int n, int max_keep, int32_t* __restrict__ keep, int32_t* num_keep) {
    const int col_blocks = (n + 63) / 64;                          // <= 256 (n <= 16384)
    __shared__ unsigned long long diag2[2][64];                   // diagonal words of block blk (and blk+1, prefetched)
    __shared__ unsigned long long remv[256];                      // removal word of every column block
    __shared__ unsigned long long part[SW_THREADS];
    __shared__ int s_list[64];
    __shared__ int s_nk, s_num;
    const int tid = threadIdx.x;
    if (tid == 0) s_num = 0;
    remv[tid] = 0ull;
    if (tid < min(64, n)) diag2[0][tid] = mask[(int64_t)tid * col_blocks];
    __syncthreads();
    for (int blk = 0; blk < col_blocks; ++blk) {
        const int base = blk * 64;
        const int cnt = min(64, n - base);
        const unsigned long long* diag = diag2[blk & 1];
        // the next block's diagonal does not depend on this block's outcome: fetch it now, one L2 round trip off the chain
        unsigned long long nd = 0ull;
        const bool pf = tid < 64 && (blk + 1 < col_blocks) && (base + 64 + tid < n);
        if (pf) nd = mask[(int64_t)(base + 64 + tid) * col_blocks + blk + 1];
        if (tid == 0) {
            unsigned long long r = remv[blk];
            int num = s_num, nk = 0;
            for (int i = 0; i < cnt; ++i) {
                const unsigned long long d = diag[i];
                const bool k = !((r >> i) & 1ull) && num < max_keep;
                if (k) { keep[num] = base + i; s_list[nk] = i; }
                num += k; nk += k;
                r |= k ? d : 0ull;
            }
            s_nk = nk; s_num = num;
        }
        __syncthreads();
        if (s_num >= max_keep) break;
        if (pf) diag2[(blk + 1) & 1][tid] = nd;
        const int nk = s_nk;
        // OR the kept rows into the removal words of the later column blocks
        for (int c0 = blk + 1; c0 < col_blocks; c0 += 64) {
            const int c = c0 + (tid & 63), qtr = tid >> 6;
            unsigned long long o = 0ull;
            if (c < col_blocks) {
                for (int u0 = qtr; u0 < nk; u0 += 32) {            // this quarter's rows: u0, u0+4, ... ; 8 loads in flight
                    unsigned long long v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int k = u0 + 4 * u;
                        v[u] = (k < nk) ? mask[(int64_t)(base + s_list[k]) * col_blocks + c] : 0ull;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) o |= v[u];
                }
            }
            part[tid] = o;
            __syncthreads();
            if (tid < 64 && c < col_blocks) remv[c] |= part[tid] | part[tid + 64] | part[tid + 128] | part[tid + 192];
            __syncthreads();
        }
    }
    if (tid == 0) *num_keep = s_num;
}

__global__ void __launch_bounds__(64)
k_nms_mask_if_overflow(const float* __restrict__ boxes, int n, float thresh, unsigned long long* __restrict__ mask,
                       const unsigned int* __restrict__ count, unsigned int cap) {
    if (*count <= cap) return;
    const int rb = blockIdx.y, cb = blockIdx.x;
    const int col_blocks = (n + 63) / 64;
    __shared__ float sb[64 * 7];
    const int ncol = min(64, n - cb * 64), nrow = min(64, n - rb * 64);
    if (threadIdx.x < ncol) {
        const float* s = boxes + (int64_t)(cb * 64 + threadIdx.x) * 7;
#pragma unroll
        for (int k = 0; k < 7; ++k) sb[threadIdx.x * 7 + k] = s[k];
    }
    __syncthreads();
    if (threadIdx.x >= nrow || cb < rb) return;
    const int i = rb * 64 + threadIdx.x;
    float a[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) a[k] = boxes[(int64_t)i * 7 + k];
    unsigned long long bits = 0ull;
    const int start = (rb == cb) ? threadIdx.x + 1 : 0;
    for (int j = start; j < ncol; ++j)
        if (rot_iou_bev(a, sb + j * 7) > thresh) bits |= 1ull << j;
    mask[(int64_t)i * col_blocks + cb] = bits;
}

extern "C" int64_t insmos_nms_pair_capacity(int32_t n) {
    // ~20 spatial neighbours per box in a dense BEV map; 128 per box is generous and keeps the list at 4 MB for 4096 boxes
    return n <= 0 ? 0 : (int64_t)n * 128;
}

extern "C" int insmos_nms_rotated(const float* boxes, int32_t n, float thresh, int32_t max_keep,
                                  unsigned long long* mask, int32_t* keep, int32_t* num_keep, void* stream) {
    if ((n > 0 && !boxes) || !mask || !keep || !num_keep || n < 0 || max_keep <= 0) return INSMOS_ERR_INVALID_ARG;
    if (n > 16384) return INSMOS_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { INSMOS_CHECK_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int32_t), st)); return INSMOS_OK; }
    const int cb = (n + 63) / 64;
    k_nms_mask<<<dim3(cb, cb), 64, 0, st>>>(boxes, n, thresh, mask);
    INSMOS_CHECK_LAUNCH("k_nms_mask");
    k_nms_sweep<<<1, SW_THREADS, 0, st>>>(mask, n, max_keep, keep, num_keep);
    INSMOS_CHECK_LAUNCH("k_nms_sweep");
    return INSMOS_OK;
}

// same result as insmos_nms_rotated through the pair list: pairs = uint2 [pair_cap] scratch, count = uint32 [1] scratch
extern "C" int insmos_nms_rotated_pairs(const float* boxes, int32_t n, float thresh, int32_t max_keep,
                                        unsigned long long* mask, void* pairs, int64_t pair_cap, uint32_t* count,
                                        int32_t* keep, int32_t* num_keep, void* stream) {
    if ((n > 0 && !boxes) || !mask || !keep || !num_keep || n < 0 || max_keep <= 0) return INSMOS_ERR_INVALID_ARG;
    if (n > 16384) return INSMOS_ERR_UNSUPPORTED;
    if (thresh < 0.0f || !pairs || !count || pair_cap <= 0 || pair_cap > 0x7fffffffll)
        return insmos_nms_rotated(boxes, n, thresh, max_keep, mask, keep, num_keep, stream);
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { INSMOS_CHECK_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int32_t), st)); return INSMOS_OK; }
    const int cb = (n + 63) / 64;
    INSMOS_CHECK_CUDA(cudaMemsetAsync(mask, 0, sizeof(unsigned long long) * (size_t)n * cb, st));
    INSMOS_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(uint32_t), st));
    k_nms_pairs<<<dim3(cb, cb), 64, 0, st>>>(boxes, n, reinterpret_cast<uint2*>(pairs), (unsigned)pair_cap, count);
    INSMOS_CHECK_LAUNCH("k_nms_pairs");
    const int64_t blocks = ceil_div64(pair_cap, 128) < 148 * 8 ? ceil_div64(pair_cap, 128) : 148 * 8;
    k_nms_pair_iou<<<(unsigned)blocks, 128, 0, st>>>(boxes, n, thresh, reinterpret_cast<const uint2*>(pairs), (unsigned)pair_cap, count, mask);
    INSMOS_CHECK_LAUNCH("k_nms_pair_iou");
    k_nms_mask_if_overflow<<<dim3(cb, cb), 64, 0, st>>>(boxes, n, thresh, mask, count, (unsigned)pair_cap);
    INSMOS_CHECK_LAUNCH("k_nms_mask_if_overflow");
    k_nms_sweep<<<1, SW_THREADS, 0, st>>>(mask, n, max_keep, keep, num_keep);
    INSMOS_CHECK_LAUNCH("k_nms_sweep");
    return INSMOS_OK;
}

// ------------------------------------------------------------------------------------------------
__global__ void k_boxes_to_voxel(const float* __restrict__ b7, const int32_t* __restrict__ labels, int nb,
                                 float x0, float y0, float z0, float vx, float vy, float vz, float stride,
                                 float* __restrict__ b8) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const float* s = b7 + i * 7; float* o = b8 + i * 8;
    // (x - min) / voxel / stride, three separate fp32 roundings (spconv_unet.py:324-329)
    o[0] = __fdiv_rn(__fdiv_rn(__fsub_rn(s[0], x0), vx), stride);
    o[1] = __fdiv_rn(__fdiv_rn(__fsub_rn(s[1], y0), vy), stride);
    o[2] = __fdiv_rn(__fdiv_rn(__fsub_rn(s[2], z0), vz), stride);
    o[3] = __fdiv_rn(__fdiv_rn(s[3], vx), stride);
    o[4] = __fdiv_rn(__fdiv_rn(s[4], vy), stride);
    o[5] = __fdiv_rn(__fdiv_rn(s[5], vz), stride);
    o[6] = s[6];
    o[7] = (float)labels[i];
}
extern "C" int insmos_boxes_to_voxel_units(const float* boxes7, const int32_t* labels, int32_t nb,
                                           const float* range_min, const float* vsize, float stride,
                                           float* boxes8, void* stream) {
    if (!boxes7 || !labels || !range_min || !vsize || !boxes8 || nb < 0) return INSMOS_ERR_INVALID_ARG;
    if (nb == 0) return INSMOS_OK;
    k_boxes_to_voxel<<<(nb + 127) / 128, 128, 0, (cudaStream_t)stream>>>(boxes7, labels, nb, range_min[0], range_min[1],
                                                                         range_min[2], vsize[0], vsize[1], vsize[2],
                                                                         stride, boxes8);
    INSMOS_CHECK_LAUNCH("k_boxes_to_voxel");
    return INSMOS_OK;
}

// ------------------------------------------------------------------------------------------------
// Array_Index.find_features_by_bbox_with_yaw semantics (Array_Index.cpp:14-79), per box:
//   first = smallest voxel index inside the box; voxel j is marked iff it is inside AND
//   (j == first OR j lies within +-extent (full extents, axis aligned) of voxel `first`).
// Two passes, thread per voxel, boxes staged in shared memory in chunks.
#define MB_CHUNK 64            // boxes per block: the chunks are a grid dimension (ncu: 13 % warps active with the chunk loop inside a thread)
struct BoxS { float cx, cy, cz, ex, ey, ez, c, s; int label; };

__device__ __forceinline__ void load_boxes(const float* __restrict__ b8, int nb, int b0, float mult, BoxS* sb) {
    for (int i = threadIdx.x; i < MB_CHUNK; i += blockDim.x) {
        const int b = b0 + i;
        if (b < nb) {
            const float* p = b8 + (int64_t)b * 8;
            BoxS q;
            q.cx = p[0] * mult; q.cy = p[1] * mult; q.cz = p[2] * mult;
            q.ex = p[3] * mult; q.ey = p[4] * mult; q.ez = p[5] * mult;
            q.c = cosf(p[6]); q.s = sinf(p[6]);
            q.label = (int)p[7];
            sb[i] = q;
        }
    }
}
__device__ __forceinline__ bool vox_inside(const BoxS& q, int x, int y, int z) {
    const float dx = (float)x - q.cx, dy = (float)y - q.cy, dz = (float)z - q.cz;
    const float rx = dx * q.c + dy * q.s;
    const float ry = -dx * q.s + dy * q.c;
    return rx <= q.ex / 2 && rx >= -q.ex / 2 && ry <= q.ey / 2 && ry >= -q.ey / 2 && dz <= q.ez / 2 && dz >= -q.ez / 2;
}

__global__ void __launch_bounds__(256)
k_member_first(const int32_t* __restrict__ coords, int64_t n, const float* __restrict__ b8, int nb, float mult,
               int32_t* first_hit) {
    __shared__ BoxS sb[MB_CHUNK];
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int x = 0, y = 0, z = 0;
    if (j < n) { const int32_t* p = coords + j * 4; z = p[1]; y = p[2]; x = p[3]; }
    {
        const int b0 = blockIdx.y * MB_CHUNK;
        load_boxes(b8, nb, b0, mult, sb);
        __syncthreads();
        const int m = min(MB_CHUNK, nb - b0);
        if (j < n)
            for (int i = 0; i < m; ++i)
                if (vox_inside(sb[i], x, y, z)) atomicMin(&first_hit[b0 + i], (int)j);
    }
}

__global__ void __launch_bounds__(256)
k_member_mark(const int32_t* __restrict__ coords, int64_t n, const float* __restrict__ b8, int nb, float mult,
              const int32_t* __restrict__ first_hit, float* __restrict__ out, int out_stride) {
    __shared__ BoxS sb[MB_CHUNK];
    __shared__ int sfx[MB_CHUNK], sfy[MB_CHUNK], sfz[MB_CHUNK], sfj[MB_CHUNK];
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int x = 0, y = 0, z = 0;
    if (j < n) { const int32_t* p = coords + j * 4; z = p[1]; y = p[2]; x = p[3]; }
    {
        const int b0 = blockIdx.y * MB_CHUNK;
        load_boxes(b8, nb, b0, mult, sb);
        for (int i = threadIdx.x; i < MB_CHUNK; i += blockDim.x) {
            const int b = b0 + i;
            int f = (b < nb) ? first_hit[b] : INT_MAX;
            sfj[i] = f;
            if (f != INT_MAX) { const int32_t* p = coords + (int64_t)f * 4; sfz[i] = p[1]; sfy[i] = p[2]; sfx[i] = p[3]; }
        }
        __syncthreads();
        const int m = min(MB_CHUNK, nb - b0);
        if (j < n)
            for (int i = 0; i < m; ++i) {
                const int f = sfj[i];
                if (f == INT_MAX || (int)j < f) continue;
                const BoxS& q = sb[i];
                if (q.label <= 0) continue;
                if ((int)j != f) {
                    // pruning window around the first hit: int coordinate vs (int + float extent), as in the reference
                    if ((float)x > (float)sfx[i] + q.ex || (float)x < (float)sfx[i] - q.ex ||
                        (float)y > (float)sfy[i] + q.ey || (float)y < (float)sfy[i] - q.ey ||
                        (float)z > (float)sfz[i] + q.ez || (float)z < (float)sfz[i] - q.ez)
                        continue;
                }
                if (vox_inside(q, x, y, z)) out[j * out_stride + (q.label - 1)] = 1.0f;
            }
    }
}

__global__ void k_fill_i32(int32_t* p, int n, int32_t v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

extern "C" int insmos_box_membership(const int32_t* coords, int64_t n, const float* boxes8, int32_t nb, float mult,
                                     float* out, int32_t out_stride, int32_t* first_hit, void* stream) {
    if (!coords || !out || n < 0 || nb < 0 || out_stride <= 0) return INSMOS_ERR_INVALID_ARG;
    if (nb == 0 || n == 0) return INSMOS_OK;
    if (!boxes8 || !first_hit) return INSMOS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    k_fill_i32<<<(nb + 255) / 256, 256, 0, st>>>(first_hit, nb, INT_MAX);
    INSMOS_CHECK_LAUNCH("k_fill_i32");
    const unsigned nblk = (unsigned)ceil_div64(n, 256);
    const dim3 grid(nblk, (unsigned)((nb + MB_CHUNK - 1) / MB_CHUNK));
    k_member_first<<<grid, 256, 0, st>>>(coords, n, boxes8, nb, mult, first_hit);
    INSMOS_CHECK_LAUNCH("k_member_first");
    k_member_mark<<<grid, 256, 0, st>>>(coords, n, boxes8, nb, mult, first_hit, out, out_stride);
    INSMOS_CHECK_LAUNCH("k_member_mark");
    return INSMOS_OK;
}
