// Input staging and output labelling around the forward path.
//
// N1  scripts/predict_mos.py:114-159,161-166,174-179 (DemoDataset.__getitem__): every past scan is moved into the frame
//     of the newest scan with inv(to_pose) @ from_pose (float64 on the host in the reference: numpy hstack / @ / .T),
//     truncated to float32, stamped with t_i = round((i - n + 1) * dt, 3) and concatenated oldest -> newest.
//     Here: the raw scans arrive as one [sum N, 4] float32 buffer (one pinned H2D copy), the n 4x4 float64 transforms
//     (n * 128 bytes, computed on the host exactly as the reference does) and the n timestamps ride along; one kernel
//     writes the [sum N, 5] float32 tensor the model consumes.  The product T @ (x,y,z,1) is evaluated in float64 and
//     rounded once to float32, as numpy does (the summation order of the 4 terms is the dgemm order k = 0..3).
// N2  scripts/predict_mos.py:440-454,279-283: logits -> ignored classes to -inf -> softmax -> confidence[:, 1:],
//     argmax -> learning_map_inv -> int32 label.  One kernel; the two small outputs leave by asynchronous D2H copies.
#include "common.cuh"

__global__ void k_stage_scans(const float* __restrict__ raw, const int64_t* __restrict__ offsets, int n_scans,
                              const double* __restrict__ transforms, const float* __restrict__ stamps,
                              int apply_transform, float* __restrict__ out, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int s = 0;                                                    // scan of point i: n_scans <= ~32, offsets are uniform loads
    while (s + 1 < n_scans && i >= offsets[s + 1]) ++s;
    const float4 p = __ldg(reinterpret_cast<const float4*>(raw) + i);
    float x = p.x, y = p.y, z = p.z;
    if (apply_transform) {
        const double* T = transforms + s * 16;
        const double dx = (double)p.x, dy = (double)p.y, dz = (double)p.z;
        // row r: ((T[r][0]*x + T[r][1]*y) + T[r][2]*z) + T[r][3]*1, products and sums rounded separately (-fmad=false)
        x = (float)(((T[0] * dx + T[1] * dy) + T[2] * dz) + T[3]);
        y = (float)(((T[4] * dx + T[5] * dy) + T[6] * dz) + T[7]);
        z = (float)(((T[8] * dx + T[9] * dy) + T[10] * dz) + T[11]);
    }
    float* o = out + i * 5;
    o[0] = x; o[1] = y; o[2] = z; o[3] = p.w; o[4] = stamps[s];
}

extern "C" int insmos_stage_scans(const float* raw_xyzi, const int64_t* scan_offsets, int32_t n_scans,
                                  const double* transforms, const float* timestamps, int32_t apply_transform,
                                  float* out_xyzit, int64_t total_points, void* stream) {
    if (n_scans <= 0 || total_points < 0 || !scan_offsets || !timestamps || (apply_transform && !transforms)) return INSMOS_ERR_INVALID_ARG;
    if (total_points == 0) return INSMOS_OK;
    if (!raw_xyzi || !out_xyzit) return INSMOS_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(raw_xyzi) & 15) != 0) return INSMOS_ERR_INVALID_ARG;
    k_stage_scans<<<(unsigned)ceil_div64(total_points, 256), 256, 0, (cudaStream_t)stream>>>(
        raw_xyzi, scan_offsets, n_scans, transforms, timestamps, apply_transform, out_xyzit, total_points);
    INSMOS_CHECK_LAUNCH("k_stage_scans");
    return INSMOS_OK;
}

__global__ void k_mos_labels(const float* __restrict__ logits, int64_t n, int C, unsigned ignore_mask,
                             const int32_t* __restrict__ label_map, int32_t* __restrict__ labels,
                             float* __restrict__ confidence) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v[8];
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) {
        v[c] = ((ignore_mask >> c) & 1u) ? -INFINITY : __ldg(logits + i * C + c);
        m = fmaxf(m, v[c]);
    }
    float sum = 0.f;
    for (int c = 0; c < C; ++c) { v[c] = expf(v[c] - m); sum += v[c]; }
    int best = 0;
    float bp = -1.f;
    for (int c = 0; c < C; ++c) {
        const float pr = v[c] / sum;
        if (pr > bp) { bp = pr; best = c; }                        // first maximum wins (torch.argmax)
        if (c >= 1 && confidence) confidence[i * (C - 1) + (c - 1)] = pr;
    }
    labels[i] = label_map ? __ldg(label_map + best) : best;
}

extern "C" int insmos_mos_labels(const float* logits, int64_t n, int32_t n_class, uint32_t ignore_mask,
                                 const int32_t* label_map, int32_t* labels, float* confidence, void* stream) {
    if (n < 0 || n_class < 2 || n_class > 8) return INSMOS_ERR_INVALID_ARG;
    if ((ignore_mask & ((1u << n_class) - 1u)) == ((1u << n_class) - 1u)) return INSMOS_ERR_INVALID_ARG;   // nothing left to pick
    if (n == 0) return INSMOS_OK;
    if (!labels || !logits) return INSMOS_ERR_INVALID_ARG;
    k_mos_labels<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(logits, n, n_class, ignore_mask, label_map, labels, confidence);
    INSMOS_CHECK_LAUNCH("k_mos_labels");
    return INSMOS_OK;
}
