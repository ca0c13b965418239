// Training-side kernels (BASELINE config 5): weight gradient of the sparse convolution over the
// tiled rule book, column moments / backward reductions of train-mode BatchNorm, scatter-add (backward of the row gathers),
// CenterHead target assignment (center_head.py:126-249) and a fused Adam step over a flat parameter buffer.
// The DATA gradient of a sparse convolution needs no kernel of its own: it is the forward kernel run over the transposed
// rule book with transposed weights (insmos_b200/autograd.py).
#include "common.cuh"
#include <math.h>

// ------------------------------------------------------------------------------------------------
// wgrad:  dW[k][ci][co] = sum over the pairs (i, o) of bucket k of  x[i][ci] * dy[o][co]
//
// grid = (K, S, ceil(Cin / CIB)); a block owns offset k, the tiles t = s, s + S, ... of slice s and CIB input channels.
// Warp w of the block walks the tiles  s + S * (w + 8 j).  A lane owns FOUR consecutive output channels (one float4 of the dy
// row) and CIB input channels: acc[CIB][4]; LPG = CoutP / 4 lanes cover a row, so a warp takes G = 32 / LPG pairs side by side
// and U of them in flight per lane: per pair and lane 1 float4 of dy + CIB / 4 float4 of x (the lanes of a group read the same
// x row: one sector each) feed 4 * CIB FFMAs.  First version (one channel per lane, profiles/r02_train_calls.jsonl): 5 loads per
// 16 FFMAs and one pair per warp at Cout = 32 -- 1.7-5 TFLOP/s on the MotionNet layers.
// The lane groups are then summed by shuffles, the warps through shared memory (fixed order), and the block writes ITS
// partial matrix partial[s][k][ci][co]; k_wgrad_reduce adds the S partials in slice order -- no atomics, the result does not
// depend on scheduling.
#define WG_CIB 16
#define WG_WARPS 8
// CIB = input channels per block: 16, or 8 for the 8-channel layers / 4 for 1..4 channels (no padded accumulators; more pairs in flight)
template <int U, int CIB>
__global__ void __launch_bounds__(WG_WARPS * 32)
k_spconv_wgrad(const float* __restrict__ x, const float* __restrict__ dy, const uint16_t* __restrict__ seg,
               const uint32_t* __restrict__ entries, int TM, int K, int Cin, int Cout, int LPG, int64_t n_tiles, int S,
               int aligned, float* __restrict__ partial) {
    __shared__ float red[WG_WARPS][CIB][33];
    const int k = blockIdx.x, s = blockIdx.y, ci0 = blockIdx.z * CIB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int G = 32 / LPG, g = lane / LPG, cq = lane - g * LPG;       // pair group, channel quad of this lane
    const int c0 = 4 * cq;
    float acc[CIB][4];
#pragma unroll
    for (int a = 0; a < CIB; ++a)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[a][q] = 0.0f;
    const bool xvec = (aligned & 1) && (Cin & 3) == 0 && ci0 + CIB <= Cin;      // x rows: 16-byte aligned float4 chunks
    const bool dvec = (aligned & 2) && (Cout & 3) == 0;                             // dy rows likewise
    for (int64_t tile = s + (int64_t)S * warp; tile < n_tiles; tile += (int64_t)S * WG_WARPS) {
        const uint16_t* tseg = seg + tile * (K + 1);
        const int start = tseg[k], n = (int)tseg[k + 1] - start;
        const uint32_t* tent = entries + tile * (int64_t)TM * K + start;
        for (int p0 = 0; p0 < n; p0 += G * U) {
            uint32_t e[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { const int p = p0 + g + G * u; e[u] = p < n ? __ldg(tent + p) : 0xffffffffu; }
            float d[U][4], xv[U][CIB];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool ok = e[u] != 0xffffffffu;
                const int64_t i = ok ? (int64_t)(e[u] & INSMOS_ROW_MASK) : 0, o = ok ? tile * TM + (e[u] >> INSMOS_ROW_BITS) : 0;
                const float* dr = dy + o * Cout + c0;
                if (dvec) {
                    const float4 v = (ok && c0 < Cout) ? __ldg(reinterpret_cast<const float4*>(dr)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    d[u][0] = v.x; d[u][1] = v.y; d[u][2] = v.z; d[u][3] = v.w;
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) d[u][q] = (ok && c0 + q < Cout) ? __ldg(dr + q) : 0.0f;
                }
                const float* xr = x + i * Cin + ci0;
                if (xvec) {
#pragma unroll
                    for (int q = 0; q < CIB / 4; ++q) {
                        const float4 v = ok ? __ldg(reinterpret_cast<const float4*>(xr) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                        xv[u][4 * q] = v.x; xv[u][4 * q + 1] = v.y; xv[u][4 * q + 2] = v.z; xv[u][4 * q + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < CIB; ++a) xv[u][a] = (ok && ci0 + a < Cin) ? __ldg(xr + a) : 0.0f;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)                                 // fixed order: the sum does not depend on scheduling
#pragma unroll
                for (int a = 0; a < CIB; ++a)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[a][q] = __fmaf_rn(xv[u][a], d[u][q], acc[a][q]);
        }
    }
    // lane groups -> group 0 (fixed shuffle tree), then per channel-of-the-quad q: warps -> shared memory -> fixed-order sum
    float* dst = partial + (((size_t)s * K + k) * Cin) * Cout;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int a = 0; a < CIB; ++a) {
            float v = acc[a][q];
            for (int off = 16; off >= LPG; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
            if (g == 0) red[warp][a][cq] = v;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < CIB * 32; idx += blockDim.x) {
            const int a = idx >> 5, c = idx & 31, co = 4 * c + q;
            if (ci0 + a >= Cin || co >= Cout || c >= LPG) continue;
            float v = 0.0f;
#pragma unroll
            for (int w = 0; w < WG_WARPS; ++w) v += red[w][a][c];
            dst[(size_t)(ci0 + a) * Cout + co] = v;
        }
        __syncthreads();
    }
}

__global__ void k_wgrad_reduce(const float* __restrict__ partial, int S, int64_t n, float* __restrict__ dw) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = 0.0f;
    for (int s = 0; s < S; ++s) v += partial[(size_t)s * n + i];
    dw[i] = v;
}

extern "C" int32_t insmos_sparse_conv_wgrad_slices(int64_t n_out, int32_t TM, int32_t K, int32_t Cin) {
    const int64_t n_tiles = ceil_div64(n_out > 0 ? n_out : 1, TM);
    const int cib = Cin <= 4 ? 4 : (Cin <= 8 ? 8 : WG_CIB);
    const int nz = (Cin + cib - 1) / cib;
    int64_t S = ceil_div64(1184, (int64_t)K * nz);                     // ~8 blocks per SM
    if (S > 64) S = 64;
    if (S > n_tiles) S = n_tiles;
    if (S < 1) S = 1;
    return (int32_t)S;
}

extern "C" int insmos_sparse_conv_wgrad(const float* in, int64_t n_in, int32_t Cin, const float* dout, int64_t n_out,
                                        int32_t Cout, const uint16_t* seg, const uint32_t* entries, int32_t TM, int32_t K,
                                        float* partial, int32_t S, float* dweight, void* stream) {
    if (!dweight || !partial || !seg || !entries || Cin <= 0 || Cout <= 0 || K <= 0 || TM <= 0 || n_in < 0 || n_out < 0 ||
        (n_in > 0 && !in) || (n_out > 0 && !dout))
        return INSMOS_ERR_INVALID_ARG;
    if (S != insmos_sparse_conv_wgrad_slices(n_out, TM, K, Cin)) return INSMOS_ERR_INVALID_ARG;
    if (Cout > 128 || K > 65535) return INSMOS_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = (int64_t)K * Cin * Cout;
    if (n_out == 0 || n_in == 0) {
        INSMOS_CHECK_CUDA(cudaMemsetAsync(dweight, 0, sizeof(float) * n, st));
        return INSMOS_OK;
    }
    int LPG = 1;                                                          // lanes per row: next power of two >= ceil(Cout / 4)
    while (LPG * 4 < Cout) LPG <<= 1;
    const int64_t n_tiles = ceil_div64(n_out, TM);
    const int cib = Cin <= 4 ? 4 : (Cin <= 8 ? 8 : WG_CIB);
    const dim3 grid((unsigned)K, (unsigned)S, (unsigned)((Cin + cib - 1) / cib));
    const int aligned = ((reinterpret_cast<uintptr_t>(in) & 15) == 0 ? 1 : 0) | ((reinterpret_cast<uintptr_t>(dout) & 15) == 0 ? 2 : 0);
    if (cib == 4) k_spconv_wgrad<4, 4><<<grid, WG_WARPS * 32, 0, st>>>(in, dout, seg, entries, TM, K, Cin, Cout, LPG, n_tiles, S, aligned, partial);
    else if (cib == 8) k_spconv_wgrad<4, 8><<<grid, WG_WARPS * 32, 0, st>>>(in, dout, seg, entries, TM, K, Cin, Cout, LPG, n_tiles, S, aligned, partial);
    else k_spconv_wgrad<2, WG_CIB><<<grid, WG_WARPS * 32, 0, st>>>(in, dout, seg, entries, TM, K, Cin, Cout, LPG, n_tiles, S, aligned, partial);
    INSMOS_CHECK_LAUNCH("k_spconv_wgrad");
    k_wgrad_reduce<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(partial, S, n, dweight);
    INSMOS_CHECK_LAUNCH("k_wgrad_reduce");
    return INSMOS_OK;
}

// ------------------------------------------------------------------------------------------------
// Column reductions over x [n, C] (C <= 1024), the building block of train-mode BatchNorm (forward: mean, then the centred
// second moment -- two passes, the variance is never formed as E[x^2] - E[x]^2; backward: sum(dy) and sum(dy * xhat)).
//   mode 0: out0[c] = sum_r a[r][c]
//   mode 1: out0[c] = sum_r (a[r][c] - mean[c])^2
//   mode 2: g = gate ? (gate[r][c] > 0 ? a : 0) : a;  out0[c] = sum_r g,  out1[c] = sum_r g * (b[r][c] - mean[c]) * invstd[c]
//   mode 3: out0[c] = sum_r a[r][c],  out1[c] = sum_r a[r][c]^2   (one pass; both sums are exact products accumulated in fp64,
//           so var = out1/n - mean^2 formed in fp64 loses nothing an fp32 two-pass variance would keep)
// Block partials are accumulated in double; a block adds its column sums to the double outputs with atomicAdd (the order
// of the ~600 block contributions varies, in double that is invisible after rounding to fp32).
#define CM_THREADS 256
__global__ void __launch_bounds__(CM_THREADS)
k_column_moments(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gate,
                 const float* __restrict__ mean, const float* __restrict__ invstd, int64_t n, int C, int mode,
                 int rows_per_block, double* __restrict__ out0, double* __restrict__ out1) {
    extern __shared__ double sh[];                                      // [2][RL][C]
    const int RL = CM_THREADS / C > 0 ? CM_THREADS / C : 1;             // row lanes (C <= 256), else columns are looped
    const int CS = C <= CM_THREADS ? C : CM_THREADS;                    // column stride of the shared partials
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = (r0 + rows_per_block < n) ? r0 + rows_per_block : n;
    for (int cbase = 0; cbase < C; cbase += CM_THREADS) {
        const int c = cbase + (C <= CM_THREADS ? threadIdx.x % C : threadIdx.x);
        const int rl = C <= CM_THREADS ? threadIdx.x / C : 0;
        double s0 = 0.0, s1 = 0.0;
        if (c < C && rl < RL) {
            const float mu = (mode == 1 || mode == 2) ? mean[c] : 0.0f;
            const float is = (mode == 2) ? invstd[c] : 0.0f;
            for (int64_t r = r0 + rl; r < r1; r += RL) {
                const float v = a[r * C + c];
                if (mode == 0) s0 += (double)v;
                else if (mode == 3) { s0 += (double)v; s1 += (double)v * (double)v; }
                else if (mode == 1) { const float d = v - mu; s0 += (double)d * (double)d; }
                else {
                    const float gv = (gate && !(gate[r * C + c] > 0.0f)) ? 0.0f : v;
                    s0 += (double)gv;
                    s1 += (double)gv * (double)((b[r * C + c] - mu) * is);
                }
            }
        }
        if (c < C && rl < RL) { sh[rl * CS + (c - cbase)] = s0; sh[(RL + rl) * CS + (c - cbase)] = s1; }
        __syncthreads();
        if (rl == 0 && c < C) {
            double t0 = 0.0, t1 = 0.0;
            for (int q = 0; q < RL; ++q) { t0 += sh[q * CS + (c - cbase)]; t1 += sh[(RL + q) * CS + (c - cbase)]; }
            atomicAdd(out0 + c, t0);
            if (mode >= 2) atomicAdd(out1 + c, t1);
        }
        __syncthreads();
    }
}

extern "C" int insmos_column_moments(const float* a, const float* b, const float* gate, const float* mean, const float* invstd,
                                     int64_t n, int32_t C, int32_t mode, double* out0, double* out1, void* stream) {
    if (!out0 || C <= 0 || C > 1024 || n < 0 || mode < 0 || mode > 3 || (n > 0 && !a) || ((mode == 1 || mode == 2) && !mean) ||
        (mode == 2 && (!b || !invstd)) || (mode >= 2 && !out1))
        return INSMOS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    INSMOS_CHECK_CUDA(cudaMemsetAsync(out0, 0, sizeof(double) * C, st));
    if (mode >= 2) INSMOS_CHECK_CUDA(cudaMemsetAsync(out1, 0, sizeof(double) * C, st));
    if (n == 0) return INSMOS_OK;
    int64_t blocks = 148 * 4;
    int64_t rpb = ceil_div64(n, blocks);
    if (rpb < 64) rpb = 64;
    blocks = ceil_div64(n, rpb);
    const int Cc = C <= CM_THREADS ? C : CM_THREADS;
    const int RL = CM_THREADS / Cc > 0 ? CM_THREADS / Cc : 1;
    const size_t smem = sizeof(double) * 2 * RL * Cc;
    k_column_moments<<<(unsigned)blocks, CM_THREADS, smem, st>>>(a, b, gate, mean, invstd, n, C, mode, (int)rpb, out0, out1);
    INSMOS_CHECK_LAUNCH("k_column_moments");
    return INSMOS_OK;
}

// train-mode BatchNorm, per-channel constants from the fp64 column sums of insmos_column_moments(mode 3): batch mean, biased
// variance -> invstd, the fused affine (scale, shift) for insmos_affine_act, and the running-statistics update of
// nn.BatchNorm1d (momentum m: running = (1 - m) * running + m * batch, the variance unbiased by n / (n - 1)).
__global__ void k_bn_train_finalize(const double* __restrict__ sum, const double* __restrict__ sumsq, double n, int C,
                                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                                    float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ scale,
                                    float* __restrict__ shift, float* __restrict__ running_mean, float* __restrict__ running_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double mu = sum[c] / n;
    double var = sumsq[c] / n - mu * mu;
    if (var < 0.0) var = 0.0;
    const float muf = (float)mu;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = (gamma ? gamma[c] : 1.0f) * is;
    mean[c] = muf; invstd[c] = is; scale[c] = sc;
    shift[c] = (beta ? beta[c] : 0.0f) - muf * sc;
    if (running_mean) running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * muf;
    if (running_var) running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)(var * (n / (n > 1.0 ? n - 1.0 : 1.0)));
}

extern "C" int insmos_bn_train_finalize(const double* sum, const double* sumsq, int64_t n, int32_t C, const float* gamma,
                                        const float* beta, float eps, float momentum, float* mean, float* invstd, float* scale,
                                        float* shift, float* running_mean, float* running_var, void* stream) {
    if (!sum || !sumsq || n <= 0 || C <= 0 || !mean || !invstd || !scale || !shift) return INSMOS_ERR_INVALID_ARG;
    k_bn_train_finalize<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sum, sumsq, (double)n, C, gamma, beta, eps, momentum, mean, invstd,
                                                                           scale, shift, running_mean, running_var);
    INSMOS_CHECK_LAUNCH("k_bn_train_finalize");
    return INSMOS_OK;
}

// train-mode BatchNorm backward, elementwise part:  dx = gamma * invstd * (g - s0/n - xhat * s1/n)  with g = dy gated by the
// fused ReLU (gate > 0), xhat = (x - mean) * invstd, s0 = sum(g), s1 = sum(g * xhat) (fp64, insmos_column_moments mode 2).
// The first C threads also write dbeta = s0 and dgamma = s1.
__global__ void k_bn_bwd_apply(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gate,
                               const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                               const double* __restrict__ s0, const double* __restrict__ s1, double inv_n, int64_t total, int C,
                               float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C) {
        if (dbeta) dbeta[i] = (float)s0[i];
        if (dgamma) dgamma[i] = (float)s1[i];
    }
    if (i >= total) return;
    const int c = (int)(i % C);
    float g = dy[i];
    if (gate && !(gate[i] > 0.0f)) g = 0.0f;
    const float is = invstd[c];
    const float xhat = (x[i] - mean[c]) * is;
    const float m0 = (float)(s0[c] * inv_n), m1 = (float)(s1[c] * inv_n);
    dx[i] = (gamma ? gamma[c] : 1.0f) * is * (g - m0 - xhat * m1);
}

extern "C" int insmos_bn_bwd_apply(const float* dy, const float* x, const float* gate, const float* mean, const float* invstd,
                                   const float* gamma, const double* s0, const double* s1, int64_t n, int32_t C, float* dx,
                                   float* dgamma, float* dbeta, void* stream) {
    if (C <= 0 || n <= 0 || !dy || !x || !dx || !mean || !invstd || !s0 || !s1) return INSMOS_ERR_INVALID_ARG;
    const int64_t total = n * C;
    k_bn_bwd_apply<<<(unsigned)ceil_div64(total > C ? total : C, 256), 256, 0, (cudaStream_t)stream>>>(
        dy, x, gate, mean, invstd, gamma, s0, s1, 1.0 / (double)n, total, C, dx, dgamma, dbeta);
    INSMOS_CHECK_LAUNCH("k_bn_bwd_apply");
    return INSMOS_OK;
}

// ------------------------------------------------------------------------------------------------
// out[idx[i], :] += src[i, :]  (idx < 0 skipped): backward of insmos_gather_rows / SparseTensor.slice /
// gather_features_by_pc_voxel_id.  `out` must be initialised by the caller.  fp32 atomics: the summation order of the
// points of one voxel varies between runs (a few ulps on the affected rows).
__global__ void k_scatter_add_rows(const float* __restrict__ src, const int32_t* __restrict__ idx, int64_t n, int C, int ldsrc,
                                   float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * C) return;
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    const int dst = idx[r];
    if (dst >= 0) atomicAdd(out + (int64_t)dst * C + c, src[r * ldsrc + c]);
}

extern "C" int insmos_scatter_add_rows(const float* src, int32_t C, int32_t ldsrc, const int32_t* idx, int64_t n, float* out, void* stream) {
    if (C <= 0 || ldsrc < C || n < 0 || (n > 0 && (!src || !idx || !out))) return INSMOS_ERR_INVALID_ARG;
    if (n == 0) return INSMOS_OK;
    k_scatter_add_rows<<<(unsigned)ceil_div64(n * C, 256), 256, 0, (cudaStream_t)stream>>>(src, idx, n, C, ldsrc, out);
    INSMOS_CHECK_LAUNCH("k_scatter_add_rows");
    return INSMOS_OK;
}

// ------------------------------------------------------------------------------------------------
// CenterHead.get_targets_single (center_head.py:171-249) for one sample: one block per ground-truth box.
// The reference walks the boxes in a Python loop (a device synchronisation per comparison); the heat map is a running
// maximum, which commutes, so the boxes are drawn in parallel with atomicMax on the bit pattern of the non-negative floats.
// Arithmetic follows the reference's dtypes: box sizes in fp32 (box / fp32 voxel / factor), gaussian_radius in fp32 with the
// same operation order and no fused multiply-adds, the centre in fp32 when the configured range is integral (fp32 box - int64
// range stays fp32 in torch; the reference's config.yaml lists integers) or in fp64 rounded to fp32 when it is not (fp32 -
// fp64 promotes), then truncated; the Gaussian in fp64 (numpy) rounded to fp32.
__device__ __forceinline__ float gaussian_radius_f32(float height, float width, float min_overlap) {
    const float b1 = __fadd_rn(height, width);
    const float c1 = __fdiv_rn(__fmul_rn(__fmul_rn(width, height), 1.0f - min_overlap), 1.0f + min_overlap);
    const float sq1 = __fsqrt_rn(__fsub_rn(__fmul_rn(b1, b1), __fmul_rn(4.0f, c1)));
    const float r1 = __fdiv_rn(__fadd_rn(b1, sq1), 2.0f);
    const float b2 = __fmul_rn(2.0f, __fadd_rn(height, width));
    const float c2 = __fmul_rn(__fmul_rn(1.0f - min_overlap, width), height);
    const float sq2 = __fsqrt_rn(__fsub_rn(__fmul_rn(b2, b2), __fmul_rn(16.0f, c2)));
    const float r2 = __fdiv_rn(__fadd_rn(b2, sq2), 2.0f);
    const float a3 = 4.0f * min_overlap;
    const float b3 = __fmul_rn(-2.0f * min_overlap, __fadd_rn(height, width));
    const float c3 = __fmul_rn(__fmul_rn(min_overlap - 1.0f, width), height);
    const float sq3 = __fsqrt_rn(__fsub_rn(__fmul_rn(b3, b3), __fmul_rn(__fmul_rn(4.0f, a3), c3)));
    const float r3 = __fdiv_rn(__fadd_rn(b3, sq3), 2.0f);
    return fminf(r1, fminf(r2, r3));
}

__global__ void k_center_targets(const float* __restrict__ gt, int n_box, int max_objs, int H, int W, int ncls, double x_min,
                                 double y_min, int range_fp64, float vx, float vy, int factor, float min_overlap, int min_radius,
                                 float* __restrict__ heatmap, float* __restrict__ anno, int64_t* __restrict__ ind,
                                 uint8_t* __restrict__ mask) {
    const int k = blockIdx.x;
    if (k >= n_box || k >= max_objs) return;
    const float* b = gt + (size_t)k * 8;
    const int cls_id = (int)(b[7] - 1.0f);                              // (label - 1).int(): truncation
    const float width = __fdiv_rn(__fdiv_rn(b[3], vx), (float)factor);
    const float length = __fdiv_rn(__fdiv_rn(b[4], vy), (float)factor);
    if (!(width > 0.0f && length > 0.0f && cls_id > -1) || cls_id >= ncls) return;
    int radius = (int)gaussian_radius_f32(length, width, min_overlap);
    if (radius < min_radius) radius = min_radius;
    const float cxf = range_fp64 ? (float)((((double)b[0] - x_min) / (double)vx) / (double)factor)
                                 : __fdiv_rn(__fdiv_rn(__fsub_rn(b[0], (float)x_min), vx), (float)factor);
    const float cyf = range_fp64 ? (float)((((double)b[1] - y_min) / (double)vy) / (double)factor)
                                 : __fdiv_rn(__fdiv_rn(__fsub_rn(b[1], (float)y_min), vy), (float)factor);
    const int x = (int)cxf, y = (int)cyf;
    if (!(0 <= x && x < W && 0 <= y && y < H)) return;
    // draw_heatmap_gaussian (center_head.py:369-399): sigma = diameter / 6, window clipped to the map
    const int diameter = 2 * radius + 1;
    const double sigma = (double)diameter / 6.0;
    const int left = min(x, radius), right = min(W - x, radius + 1), top = min(y, radius), bottom = min(H - y, radius + 1);
    const int ww = left + right, hh = top + bottom;
    int* hm = reinterpret_cast<int*>(heatmap + (size_t)cls_id * H * W);
    for (int i = threadIdx.x; i < ww * hh; i += blockDim.x) {
        const int dy = i / ww - top, dx = i % ww - left;
        const float gval = (float)exp(-(double)(dx * dx + dy * dy) / (2.0 * sigma * sigma));
        atomicMax(hm + (size_t)(y + dy) * W + (x + dx), __float_as_int(gval));
    }
    if (threadIdx.x == 0) {
        ind[k] = (int64_t)y * W + x;
        mask[k] = 1;
        float* a = anno + (size_t)k * 8;
        a[0] = cxf - (float)x; a[1] = cyf - (float)y; a[2] = b[2];
        a[3] = logf(b[3]); a[4] = logf(b[4]); a[5] = logf(b[5]);
        a[6] = sinf(b[6]); a[7] = cosf(b[6]);
    }
}

extern "C" int insmos_center_targets(const float* gt_boxes, int32_t n_box, int32_t max_objs, int32_t H, int32_t W, int32_t ncls,
                                     double x_min, double y_min, int32_t range_is_fp64, float vx, float vy, int32_t out_size_factor,
                                     float min_overlap, int32_t min_radius, float* heatmap, float* anno_boxes, int64_t* inds,
                                     uint8_t* masks, void* stream) {
    if (n_box < 0 || max_objs <= 0 || H <= 0 || W <= 0 || ncls <= 0 || !heatmap || !anno_boxes || !inds || !masks ||
        (n_box > 0 && !gt_boxes) || out_size_factor <= 0)
        return INSMOS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    INSMOS_CHECK_CUDA(cudaMemsetAsync(heatmap, 0, sizeof(float) * (size_t)ncls * H * W, st));
    INSMOS_CHECK_CUDA(cudaMemsetAsync(anno_boxes, 0, sizeof(float) * (size_t)max_objs * 8, st));
    INSMOS_CHECK_CUDA(cudaMemsetAsync(inds, 0, sizeof(int64_t) * (size_t)max_objs, st));
    INSMOS_CHECK_CUDA(cudaMemsetAsync(masks, 0, (size_t)max_objs, st));
    const int nb = n_box < max_objs ? n_box : max_objs;
    if (nb == 0) return INSMOS_OK;
    k_center_targets<<<nb, 128, 0, st>>>(gt_boxes, nb, max_objs, H, W, ncls, x_min, y_min, range_is_fp64, vx, vy, out_size_factor, min_overlap,
                                          min_radius, heatmap, anno_boxes, inds, masks);
    INSMOS_CHECK_LAUNCH("k_center_targets");
    return INSMOS_OK;
}

// ------------------------------------------------------------------------------------------------
// torch.optim.Adam (models.py:185-190: lr, weight_decay as L2 added to the gradient, betas 0.9 / 0.999, eps 1e-8) over ONE
// flat fp32 buffer holding every parameter (the gradients live in a second flat buffer, so the data-parallel all-reduce is a
// single NCCL call on it).  grad_scale folds the 1/world_size of the gradient average into the step.
__global__ void k_adam_step(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, float bc1, float bc2_sqrt,
                            float grad_scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gi = g[i] * grad_scale;
    const float pi = p[i];
    if (weight_decay != 0.0f) gi = __fmaf_rn(weight_decay, pi, gi);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
}

extern "C" int insmos_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                                float beta2, float eps, float weight_decay, int64_t step, float grad_scale, void* stream) {
    if (n < 0 || step < 1 || (n > 0 && (!param || !grad || !exp_avg || !exp_avg_sq))) return INSMOS_ERR_INVALID_ARG;
    if (n == 0) return INSMOS_OK;
    const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    k_adam_step<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                                weight_decay, bc1, bc2_sqrt, grad_scale);
    INSMOS_CHECK_LAUNCH("k_adam_step");
    return INSMOS_OK;
}
