// Instance-refinement step after the forward path: the per-point parts of
// scripts/refine.py:169-302 on the device.
//   * insmos_point_instance_ids  = Array_Index.find_point_in_instance_bbox_with_yaw (models/utils/src/Array_Index.cpp:83-149):
//       which predicted box (index + 1) of each class contains each LiDAR point, with the reference's first-hit pruning
//       window; the reference runs it on the host with 30 OpenMP threads over boxes (and races when two boxes of a class
//       share a point) -- here thread = point, boxes staged in shared memory, atomicMax reproduces the serial box order.
//   * insmos_instance_stats      = the per-instance reductions of refine.py:208-221 (points, moving points, confident points);
//   * insmos_relabel_instances   = the per-instance label overwrites of refine.py:240-257,287-292 as one gather by instance id.
// Geometry in fp32 with separately rounded products and sums (-fmad=false), as the host code computes it.
#include "common.cuh"
#include <math.h>

#define RF_CHUNK 64
struct RBox { float cx, cy, cz, ex, ey, ez, c, s; int label; };

__device__ __forceinline__ void rf_load_boxes(const float* __restrict__ b8, int nb, int b0, float out_ground, RBox* sb) {
    for (int i = threadIdx.x; i < RF_CHUNK; i += blockDim.x) {
        const int b = b0 + i;
        if (b < nb) {
            const float* p = b8 + (int64_t)b * 8;
            RBox q;
            q.cx = p[0]; q.cy = p[1]; q.cz = p[2] + out_ground;
            q.ex = p[3]; q.ey = p[4]; q.ez = p[5];
            q.c = cosf(p[6]); q.s = sinf(p[6]);
            q.label = (int)p[7];
            sb[i] = q;
        }
    }
}
__device__ __forceinline__ bool rf_inside(const RBox& q, float x, float y, float z) {
    const float dx = x - q.cx, dy = y - q.cy, dz = z - q.cz;
    const float rx = dx * q.c + dy * q.s;
    const float ry = -dx * q.s + dy * q.c;
    return rx <= q.ex / 2 && rx >= -q.ex / 2 && ry <= q.ey / 2 && ry >= -q.ey / 2 && dz <= q.ez / 2 && dz >= -q.ez / 2;
}

__global__ void __launch_bounds__(256)
k_point_first(const float* __restrict__ pts, int64_t n, int stride, const float* __restrict__ b8, int nb, float out_ground,
              int32_t* first_hit) {
    __shared__ RBox sb[RF_CHUNK];
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float x = 0.f, y = 0.f, z = 0.f;
    if (j < n) { const float* p = pts + j * stride; x = p[0]; y = p[1]; z = p[2]; }
    const int b0 = blockIdx.y * RF_CHUNK;
    rf_load_boxes(b8, nb, b0, out_ground, sb);
    __syncthreads();
    const int m = min(RF_CHUNK, nb - b0);
    if (j < n)
        for (int i = 0; i < m; ++i)
            if (rf_inside(sb[i], x, y, z)) atomicMin(&first_hit[b0 + i], (int)j);
}

__global__ void __launch_bounds__(256)
k_point_mark(const float* __restrict__ pts, int64_t n, int stride, const float* __restrict__ b8, int nb, float out_ground,
             const int32_t* __restrict__ first_hit, int32_t* __restrict__ ids, int ncls) {
    __shared__ RBox sb[RF_CHUNK];
    __shared__ float sfx[RF_CHUNK], sfy[RF_CHUNK], sfz[RF_CHUNK];
    __shared__ int sfj[RF_CHUNK];
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float x = 0.f, y = 0.f, z = 0.f;
    if (j < n) { const float* p = pts + j * stride; x = p[0]; y = p[1]; z = p[2]; }
    const int b0 = blockIdx.y * RF_CHUNK;
    rf_load_boxes(b8, nb, b0, out_ground, sb);
    for (int i = threadIdx.x; i < RF_CHUNK; i += blockDim.x) {
        const int b = b0 + i;
        const int f = (b < nb) ? first_hit[b] : INT_MAX;
        sfj[i] = f;
        if (f != INT_MAX) { const float* p = pts + (int64_t)f * stride; sfx[i] = p[0]; sfy[i] = p[1]; sfz[i] = p[2]; }
    }
    __syncthreads();
    const int m = min(RF_CHUNK, nb - b0);
    if (j >= n) return;
    for (int i = 0; i < m; ++i) {
        const int f = sfj[i];
        if (f == INT_MAX || (int)j < f) continue;
        const RBox& q = sb[i];
        if (q.label <= 0 || q.label > ncls) continue;
        if ((int)j != f) {                                   // pruning window around the first hit (Array_Index.cpp:122-125)
            if (x > sfx[i] + q.ex || x < sfx[i] - q.ex || y > sfy[i] + q.ey || y < sfy[i] - q.ey ||
                z > sfz[i] + q.ez || z < sfz[i] - q.ez)
                continue;
        }
        if (rf_inside(q, x, y, z)) atomicMax(&ids[j * ncls + (q.label - 1)], b0 + i + 1);   // serial order: the later box wins
    }
}

__global__ void k_rf_fill(int32_t* p, int64_t n, int32_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

extern "C" int insmos_point_instance_ids(const float* points, int64_t n, int32_t stride, const float* boxes8, int32_t nb,
                                         float out_ground, int32_t* ids, int32_t ncls, int32_t* first_hit, void* stream) {
    if ((n > 0 && !points) || !ids || n < 0 || nb < 0 || stride < 3 || ncls <= 0 || n > INT_MAX) return INSMOS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return INSMOS_OK;
    INSMOS_CHECK_CUDA(cudaMemsetAsync(ids, 0, sizeof(int32_t) * (size_t)n * ncls, st));
    if (nb == 0) return INSMOS_OK;
    if (!boxes8 || !first_hit) return INSMOS_ERR_INVALID_ARG;
    k_rf_fill<<<(unsigned)ceil_div64(nb, 256), 256, 0, st>>>(first_hit, nb, INT_MAX);
    INSMOS_CHECK_LAUNCH("k_rf_fill");
    const dim3 grid((unsigned)ceil_div64(n, 256), (unsigned)ceil_div64(nb, RF_CHUNK));
    k_point_first<<<grid, 256, 0, st>>>(points, n, stride, boxes8, nb, out_ground, first_hit);
    INSMOS_CHECK_LAUNCH("k_point_first");
    k_point_mark<<<grid, 256, 0, st>>>(points, n, stride, boxes8, nb, out_ground, first_hit, ids, ncls);
    INSMOS_CHECK_LAUNCH("k_point_mark");
    return INSMOS_OK;
}

// stats[b*3 + {0,1,2}] = points of instance b+1 in column `col`, of them with label == moving_label, with conf >= conf_thresh
__global__ void k_instance_stats(const int32_t* __restrict__ ids, int ncls, int col, int64_t n, const int32_t* __restrict__ labels,
                                 int moving_label, const float* __restrict__ conf, int conf_stride, float conf_thresh, int nb,
                                 int32_t* __restrict__ stats) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int id = ids[j * ncls + col];
    if (id <= 0 || id > nb) return;
    int32_t* s = stats + (int64_t)(id - 1) * 3;
    atomicAdd(s, 1);
    if (labels[j] == moving_label) atomicAdd(s + 1, 1);
    if (conf && conf[j * conf_stride] >= conf_thresh) atomicAdd(s + 2, 1);
}
extern "C" int insmos_instance_stats(const int32_t* ids, int32_t ncls, int32_t col, int64_t n, const int32_t* labels,
                                     int32_t moving_label, const float* conf, int32_t conf_stride, float conf_thresh,
                                     int32_t nb, int32_t* stats, void* stream) {
    if (!ids || !labels || !stats || n < 0 || nb < 0 || col < 0 || col >= ncls) return INSMOS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (nb > 0) INSMOS_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(int32_t) * 3 * (size_t)nb, st));
    if (n == 0 || nb == 0) return INSMOS_OK;
    k_instance_stats<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(ids, ncls, col, n, labels, moving_label, conf, conf_stride, conf_thresh, nb, stats);
    INSMOS_CHECK_LAUNCH("k_instance_stats");
    return INSMOS_OK;
}

// labels[j] = new_label[id] for the points whose instance id (column col) has new_label[id] >= 0  (new_label [nb+1], entry 0 unused)
__global__ void k_relabel_instances(const int32_t* __restrict__ ids, int ncls, int col, int64_t n,
                                    const int32_t* __restrict__ new_label, int nb, int32_t* __restrict__ labels) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int id = ids[j * ncls + col];
    if (id <= 0 || id > nb) return;
    const int v = new_label[id];
    if (v >= 0) labels[j] = v;
}
extern "C" int insmos_relabel_instances(const int32_t* ids, int32_t ncls, int32_t col, int64_t n, const int32_t* new_label,
                                        int32_t nb, int32_t* labels, void* stream) {
    if (!ids || !new_label || !labels || n < 0 || nb < 0 || col < 0 || col >= ncls) return INSMOS_ERR_INVALID_ARG;
    if (n == 0 || nb == 0) return INSMOS_OK;
    k_relabel_instances<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(ids, ncls, col, n, new_label, nb, labels);
    INSMOS_CHECK_LAUNCH("k_relabel_instances");
    return INSMOS_OK;
}
