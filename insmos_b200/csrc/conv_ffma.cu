// Sparse convolution, fp32 FFMA path with register-blocked outputs ("thread per pair").
//
// Why this exists (measurements in profiles/r01_conv_v2_sass_notes.md): on B200 the legacy mma.sync TF32 path needs
// the 3x hi/lo split for fp32 parity and then peaks at ~48 TFLOP/s effective -- below the 72 TFLOP/s of plain fp32
// FFMA -- and its per-chunk bookkeeping instructions serve only 16 pairs each.  Here one THREAD owns one rule-book pair:
// it gathers its input row (16-byte loads), multiplies it against the bucket's [Cin x COT] weight block with COT
// independent accumulators in registers (weights are warp-uniform: one broadcast 16-byte load feeds 4 FFMAs of all
// 32 lanes), and adds its COT results into the tile's accumulator row in shared memory (within one bucket every output
// row occurs once: plain read-modify-write, no atomics).  Every bookkeeping instruction is amortised over 32 pairs and
// the arithmetic is exact fp32 (no split), so parity with the oracle is at rounding-order level.
//   k_ffma_warp<COT>  Cout <= 16: one independent warp per tile, weights read through L1 (uniform __ldg).
//   k_ffma_block<COB> Cout >= 32: a block owns a super-tile (G tiles, ~128 rows) x COB output channels; the bucket's
//                     [Cin x COB] weight block is staged once in shared memory; (32-pair chunk, 32-channel slice)
//                     work items are dealt round-robin to the 8 warps.
// Each output row is written once with the fused epilogue (BatchNorm scale/shift, bias, residual, ReLU).
#include "common.cuh"
#include <stdlib.h>

__device__ __forceinline__ float ff_epilogue(float v, int c, int64_t row, int Cout, const insmos_epilogue_t& ep) {
    if (ep.scale) v = __fmaf_rn(v, __ldg(ep.scale + c), __ldg(ep.shift + c));
    if (ep.bias) v += __ldg(ep.bias + c);
    if (ep.residual) v += __ldg(ep.residual + row * Cout + c);
    if (ep.relu) v = fmaxf(v, 0.0f);
    return v;
}

struct FfArgs {
    const float* in; const float* w; const uint16_t* seg; const uint32_t* entries; float* out;
    int64_t n_out, n_tiles; int Cin, Cout, K, TM;
    insmos_epilogue_t ep;
};

// acc[COT] += x[0..Cin) . W[ci][co0 .. co0+COT)   (W row stride = wstride floats; wp points at W[0][co0])
// WLOAD: how a float4 of weights is fetched (global uniform load or shared memory)
template <int COT, bool SMEM_W>
__device__ __forceinline__ void row_times_block(const float* __restrict__ x, int Cin, const float* wp, int wstride,
                                                float (&acc)[COT]) {
    int ci = 0;
    if ((Cin & 3) == 0) {
        for (; ci < Cin; ci += 4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(x + ci));
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float* wr = wp + (size_t)(ci + u) * wstride;
#pragma unroll
                for (int c4 = 0; c4 < COT / 4; ++c4) {
                    const float4 w = SMEM_W ? *reinterpret_cast<const float4*>(wr + c4 * 4)
                                            : __ldg(reinterpret_cast<const float4*>(wr + c4 * 4));
                    acc[c4 * 4 + 0] = __fmaf_rn(av[u], w.x, acc[c4 * 4 + 0]);
                    acc[c4 * 4 + 1] = __fmaf_rn(av[u], w.y, acc[c4 * 4 + 1]);
                    acc[c4 * 4 + 2] = __fmaf_rn(av[u], w.z, acc[c4 * 4 + 2]);
                    acc[c4 * 4 + 3] = __fmaf_rn(av[u], w.w, acc[c4 * 4 + 3]);
                }
            }
        }
    } else {
        for (; ci < Cin; ++ci) {
            const float a = __ldg(x + ci);
            const float* wr = wp + (size_t)ci * wstride;
#pragma unroll
            for (int c4 = 0; c4 < COT / 4; ++c4) {
                const float4 w = SMEM_W ? *reinterpret_cast<const float4*>(wr + c4 * 4)
                                        : __ldg(reinterpret_cast<const float4*>(wr + c4 * 4));
                acc[c4 * 4 + 0] = __fmaf_rn(a, w.x, acc[c4 * 4 + 0]);
                acc[c4 * 4 + 1] = __fmaf_rn(a, w.y, acc[c4 * 4 + 1]);
                acc[c4 * 4 + 2] = __fmaf_rn(a, w.z, acc[c4 * 4 + 2]);
                acc[c4 * 4 + 3] = __fmaf_rn(a, w.w, acc[c4 * 4 + 3]);
            }
        }
    }
}

template <int COT>
__device__ __forceinline__ void add_to_tile(float* row, const float (&acc)[COT]) {
#pragma unroll
    for (int c4 = 0; c4 < COT / 4; ++c4) {
        float4* q = reinterpret_cast<float4*>(row + c4 * 4);
        float4 v = *q;
        v.x += acc[c4 * 4 + 0]; v.y += acc[c4 * 4 + 1]; v.z += acc[c4 * 4 + 2]; v.w += acc[c4 * 4 + 3];
        *q = v;
    }
}

// ------------------------------------------------------------------------------------------------
#define FW_WARPS 4
template <int COT>
__global__ void __launch_bounds__(FW_WARPS * 32)
k_ffma_warp(FfArgs p) {
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile = (int64_t)blockIdx.x * FW_WARPS + warp;
    if (tile >= p.n_tiles) return;                            // warps are independent: no block barrier below
    const int TM = p.TM, K = p.K, Cin = p.Cin, Cout = p.Cout;
    float* acc_t = sm + (size_t)warp * TM * COT;              // [TM][COT]
    int* sseg = reinterpret_cast<int*>(sm + (size_t)FW_WARPS * TM * COT) + warp * (K + 1);
    const uint16_t* tseg = p.seg + tile * (K + 1);
    for (int k = lane; k <= K; k += 32) sseg[k] = tseg[k];
    for (int i = lane; i < TM * COT; i += 32) acc_t[i] = 0.0f;
    __syncwarp();
    const uint32_t* tent = p.entries + tile * (int64_t)TM * K;
    for (int k = 0; k < K; ++k) {
        const int s0 = sseg[k], n = sseg[k + 1] - s0;
        if (n == 0) continue;
        const float* wk = p.w + (size_t)k * Cin * Cout;
        for (int c0 = 0; c0 < n; c0 += 32) {
            const int q = c0 + lane;
            const bool valid = q < n;
            const uint32_t e = valid ? __ldg(tent + s0 + q) : 0u;
            float acc[COT];
#pragma unroll
            for (int c = 0; c < COT; ++c) acc[c] = 0.0f;
            row_times_block<COT, false>(p.in + (size_t)(e & INSMOS_ROW_MASK) * Cin, Cin, wk, Cout, acc);
            if (valid) add_to_tile<COT>(acc_t + (e >> INSMOS_ROW_BITS) * COT, acc);
            __syncwarp();
        }
    }
    const int64_t row0 = tile * TM;
    const int rows = (int)((p.n_out - row0) < TM ? (p.n_out - row0) : TM);
    for (int i = lane; i < rows * COT; i += 32) {
        const int r = i / COT, c = i % COT;
        if (c < Cout) p.out[(row0 + r) * Cout + c] = ff_epilogue(acc_t[i], c, row0 + r, Cout, p.ep);
    }
}

// ------------------------------------------------------------------------------------------------
#define FB_WARPS 8
#define FB_COT 32
template <int COB>
__global__ void __launch_bounds__(FB_WARPS * 32)
k_ffma_block(FfArgs p, int G, int n_cb) {
    constexpr int SLICES = COB / FB_COT;
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t stile = blockIdx.x / n_cb;
    const int cb0 = (int)(blockIdx.x - stile * n_cb) * COB;  // first output channel of this block
    const int TM = p.TM, K = p.K, Cin = p.Cin, Cout = p.Cout;
    const int64_t tile0 = stile * G;
    const int ntile = (int)((p.n_tiles - tile0) < G ? (p.n_tiles - tile0) : G);
    float* acc_t = sm;                                        // [G*TM][COB]
    float* wbuf = acc_t + (size_t)G * TM * COB;               // [Cin][COB]
    int* ssegs = reinterpret_cast<int*>(wbuf + (size_t)Cin * COB);   // [G][K+1]
    int* pre = ssegs + G * (K + 1);                           // [G+1] prefix of the current bucket over the tiles
    for (int i = threadIdx.x; i < G * (K + 1); i += FB_WARPS * 32) {
        const int gi = i / (K + 1), k = i - gi * (K + 1);
        ssegs[i] = (gi < ntile) ? p.seg[(tile0 + gi) * (K + 1) + k] : 0;
    }
    for (int i = threadIdx.x; i < G * TM * COB; i += FB_WARPS * 32) acc_t[i] = 0.0f;
    __syncthreads();
    for (int k = 0; k < K; ++k) {
        if (threadIdx.x == 0) {
            int run = 0;
            for (int gi = 0; gi < G; ++gi) { pre[gi] = run; run += ssegs[gi * (K + 1) + k + 1] - ssegs[gi * (K + 1) + k]; }
            pre[G] = run;
        }
        __syncthreads();
        const int tot = pre[G];
        if (tot == 0) { __syncthreads(); continue; }          // uniform
        // stage W[k][:, cb0 .. cb0+COB) (zero beyond Cout)
        const float* wk = p.w + (size_t)k * Cin * Cout;
        for (int i = threadIdx.x; i < Cin * (COB / 4); i += FB_WARPS * 32) {
            const int ci = i / (COB / 4), c4 = i - ci * (COB / 4);
            const int co = cb0 + c4 * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (co + 3 < Cout) v = __ldg(reinterpret_cast<const float4*>(wk + (size_t)ci * Cout + co));
            else if (co < Cout) {
                v.x = __ldg(wk + (size_t)ci * Cout + co);
                if (co + 1 < Cout) v.y = __ldg(wk + (size_t)ci * Cout + co + 1);
                if (co + 2 < Cout) v.z = __ldg(wk + (size_t)ci * Cout + co + 2);
            }
            *reinterpret_cast<float4*>(wbuf + (size_t)ci * COB + c4 * 4) = v;
        }
        __syncthreads();
        const int nchunk = (tot + 31) >> 5;
        for (int item = warp; item < nchunk * SLICES; item += FB_WARPS) {
            const int chunk = item / SLICES, slice = item - chunk * SLICES;
            const int q = chunk * 32 + lane;
            const bool valid = q < tot;
            int gi = 0;
            if (valid) { while (gi + 1 < G && q >= pre[gi + 1]) ++gi; }
            const uint32_t e = valid ? __ldg(p.entries + (tile0 + gi) * (int64_t)TM * K + ssegs[gi * (K + 1) + k] + (q - pre[gi])) : 0u;
            float acc[FB_COT];
#pragma unroll
            for (int c = 0; c < FB_COT; ++c) acc[c] = 0.0f;
            row_times_block<FB_COT, true>(p.in + (size_t)(e & INSMOS_ROW_MASK) * Cin, Cin, wbuf + slice * FB_COT, COB, acc);
            if (valid) add_to_tile<FB_COT>(acc_t + (size_t)(gi * TM + (int)(e >> INSMOS_ROW_BITS)) * COB + slice * FB_COT, acc);
        }
        __syncthreads();                                      // bucket done: accumulator rows and wbuf reusable
    }
    const int64_t row0 = tile0 * TM;
    const int64_t left = p.n_out - row0;
    const int rows = (int)(left < (int64_t)G * TM ? left : (int64_t)G * TM);
    for (int i = threadIdx.x; i < rows * COB; i += FB_WARPS * 32) {
        const int r = i / COB, c = cb0 + (i % COB);
        if (c < Cout) p.out[(row0 + r) * Cout + c] = ff_epilogue(acc_t[i], c, row0 + r, Cout, p.ep);
    }
}

template <class Kern>
static int set_smem(Kern kern, size_t smem, insmos_smem_cfg_t& configured) {
    INSMOS_CHECK_CUDA(insmos_ensure_smem(kern, smem, configured));
    return INSMOS_OK;
}

template <int COT>
static int launch_warp(const FfArgs& a, cudaStream_t st) {
    const size_t smem = sizeof(float) * (size_t)FW_WARPS * a.TM * COT + sizeof(int) * (size_t)FW_WARPS * (a.K + 1);
    if (smem > 220 * 1024) return INSMOS_ERR_UNSUPPORTED;
    static thread_local insmos_smem_cfg_t configured;
    int rc = set_smem(k_ffma_warp<COT>, smem, configured);
    if (rc) return rc;
    k_ffma_warp<COT><<<(unsigned)ceil_div64(a.n_tiles, FW_WARPS), FW_WARPS * 32, smem, st>>>(a);
    INSMOS_CHECK_LAUNCH("k_ffma_warp");
    return INSMOS_OK;
}

template <int COB>
static int launch_block(const FfArgs& a, cudaStream_t st) {
    int G = 128 / a.TM; if (G < 1) G = 1;
    const int n_cb = (a.Cout + COB - 1) / COB;
    const size_t smem = sizeof(float) * ((size_t)G * a.TM * COB + (size_t)a.Cin * COB) + sizeof(int) * ((size_t)G * (a.K + 1) + G + 1);
    if (smem > 220 * 1024) return INSMOS_ERR_UNSUPPORTED;
    static thread_local insmos_smem_cfg_t configured;
    int rc = set_smem(k_ffma_block<COB>, smem, configured);
    if (rc) return rc;
    k_ffma_block<COB><<<(unsigned)(ceil_div64(a.n_tiles, G) * n_cb), FB_WARPS * 32, smem, st>>>(a, G, n_cb);
    INSMOS_CHECK_LAUNCH("k_ffma_block");
    return INSMOS_OK;
}

// fp32 FFMA sparse convolution; returns INSMOS_ERR_UNSUPPORTED for shapes it does not cover (caller falls back)
// ------------------------------------------------------------------------------------------------
// Single input channel (conv0p1s1 of the 4D net: 5x5x5x1 = 125 offsets, 1 -> 8 channels, minkunet.py:55-60).
// With Cin = 1 a pair is one scalar: per tile the gathered scalars are scattered into a dense [K][TM] matrix in shared
// memory (one thread per rule-book entry, all loads independent, bucket by binary search in the tile's seg), then
// out[r][:] = sum_k A[k][r] * W[k][:] is a dense exact-fp32 FFMA loop with broadcast weight reads: 6 instructions per
// (row, offset, 4 channels) instead of ~40 per pair in the thread-per-pair kernel, and no divergence on the 13 % occupancy.
template <int COUT>
__global__ void __launch_bounds__(128)
k_spconv_cin1(FfArgs p) {
    extern __shared__ __align__(16) float sm1[];
    const int K = p.K, TM = p.TM;
    float* A = sm1;                                                  // [K][TM]
    float* Wt = A + K * TM;                                          // [K][COUT]
    int* sseg = reinterpret_cast<int*>(Wt + K * COUT);               // [K+1]
    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    for (int i = tid; i < K * TM; i += 128) A[i] = 0.0f;
    for (int i = tid; i < K * COUT; i += 128) Wt[i] = __ldg(p.w + i);
    for (int k = tid; k <= K; k += 128) sseg[k] = p.seg[tile * (K + 1) + k];
    __syncthreads();
    const int tot = sseg[K];
    const uint32_t* tent = p.entries + tile * (int64_t)TM * K;
    for (int e0 = tid; e0 < tot; e0 += 128 * 8) {                   // 8 entries, then 8 scalars, in flight per thread
        uint32_t ent[8];
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) ent[u] = (e0 + 128 * u < tot) ? __ldg(tent + e0 + 128 * u) : 0u;
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (e0 + 128 * u < tot) ? __ldg(p.in + (ent[u] & INSMOS_ROW_MASK)) : 0.0f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = e0 + 128 * u;
            if (e < tot) {
                int lo = 0, hi = K;                                  // largest k with sseg[k] <= e
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sseg[mid] <= e) lo = mid; else hi = mid; }
                A[lo * TM + (int)(ent[u] >> INSMOS_ROW_BITS)] = v[u];
            }
        }
    }
    __syncthreads();
    constexpr int Q = COUT / 4;                                      // float4 groups of output channels
    const int64_t row0 = tile * TM;
    for (int i = tid; i < TM * Q; i += 128) {
        const int r = i % TM, q = i / TM;                            // consecutive threads -> consecutive rows: conflict-free A reads
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < K; ++k) {
            const float a = A[k * TM + r];
            const float4 w = *reinterpret_cast<const float4*>(Wt + k * COUT + 4 * q);
            acc.x = __fmaf_rn(a, w.x, acc.x); acc.y = __fmaf_rn(a, w.y, acc.y);
            acc.z = __fmaf_rn(a, w.z, acc.z); acc.w = __fmaf_rn(a, w.w, acc.w);
        }
        const int64_t row = row0 + r;
        if (row < p.n_out) {
            const int c = 4 * q;
            float4 o;
            o.x = ff_epilogue(acc.x, c + 0, row, COUT, p.ep); o.y = ff_epilogue(acc.y, c + 1, row, COUT, p.ep);
            o.z = ff_epilogue(acc.z, c + 2, row, COUT, p.ep); o.w = ff_epilogue(acc.w, c + 3, row, COUT, p.ep);
            *reinterpret_cast<float4*>(p.out + row * COUT + c) = o;
        }
    }
}

template <int COUT>
static int launch_cin1(const FfArgs& a, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)a.K * a.TM + (size_t)a.K * COUT) + sizeof(int) * ((size_t)a.K + 1);
    if (smem > 200 * 1024) return INSMOS_ERR_UNSUPPORTED;
    static thread_local insmos_smem_cfg_t configured;
    INSMOS_CHECK_CUDA(insmos_ensure_smem(k_spconv_cin1<COUT>, smem, configured));
    k_spconv_cin1<COUT><<<(unsigned)a.n_tiles, 128, smem, st>>>(a);
    INSMOS_CHECK_LAUNCH("k_spconv_cin1");
    return INSMOS_OK;
}

extern "C" int insmos_sparse_conv_fwd_ffma(const float* in, int64_t n_in, int32_t Cin,
                                           const float* weight, int32_t K, int32_t Cout,
                                           const uint16_t* seg, const uint32_t* entries, int32_t TM,
                                           float* out, int64_t n_out,
                                           const insmos_epilogue_t* ep_in, void* stream) {
    if ((n_in > 0 && !in) || !weight || !seg || !entries || (n_out > 0 && !out) || Cin <= 0 || Cout <= 0 || K <= 0 || n_out < 0 || n_in < 0)
        return INSMOS_ERR_INVALID_ARG;
    if (TM != 16 && TM != 32 && TM != 64 && TM != 128) return INSMOS_ERR_INVALID_ARG;
    if (n_in > (int64_t)INSMOS_ROW_MASK + 1) return INSMOS_ERR_UNSUPPORTED;
    if (Cout % 4 != 0) return INSMOS_ERR_UNSUPPORTED;       // 16-byte weight loads
    FfArgs a;
    a.in = in; a.w = weight; a.seg = seg; a.entries = entries; a.out = out;
    a.n_out = n_out; a.n_tiles = ceil_div64(n_out, TM); a.Cin = Cin; a.Cout = Cout; a.K = K; a.TM = TM;
    a.ep = insmos_epilogue_t{nullptr, nullptr, nullptr, nullptr, 0};
    if (ep_in) a.ep = *ep_in;
    if (a.ep.scale && !a.ep.shift) return INSMOS_ERR_INVALID_ARG;
    if (n_out == 0) return INSMOS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (Cin == 1 && (Cout == 8 || Cout == 16) && K >= 27 && getenv("INSMOS_NO_CIN1") == nullptr) {
        const int rc = Cout == 8 ? launch_cin1<8>(a, st) : launch_cin1<16>(a, st);
        if (rc != INSMOS_ERR_UNSUPPORTED) return rc;
    }
    if (Cout == 8) return launch_warp<8>(a, st);
    if (Cout == 16) return launch_warp<16>(a, st);
    if (Cout <= 32) return launch_block<32>(a, st);
    return launch_block<64>(a, st);
}
