// Feature kernels: sparse convolution over the tiled rule book (a3, a8, a12), 1x1 conv / linear,
// fused BN(eval)+residual+ReLU epilogues, concat, pair-sum, gathers, dense scatter.
//
// Sparse convolution is output-stationary: a tile of TM output rows is accumulated in shared
// memory, bucket by bucket (kernel offset k).  Inside one bucket every output row appears at most
// once, so accumulation is a plain read-modify-write (shared-memory fp32 atomics are CAS loops on
// sm_100a and are avoided).  Each output row is written to HBM exactly once, with the epilogue
// (folded BatchNorm, residual, ReLU) applied in registers on the way out; every feature row is
// gathered from L2, the 4-byte rule-book entries are streamed once.
//
// Two contraction paths:
//   algo 1  SIMT fp32 FFMA, any Cin/Cout (reference path, also used for Cin < 8).
//   algo 2  tensor cores: the pairs of one bucket are packed 16 at a time into the M dimension of
//           mma.sync.m16n8k8 (100% useful rows, unlike a zero-padded output-stationary tile), the
//           [Cin x Cout] weight slice of offset k is the B operand.  fp32 parity (logits within
//           1e-3) rules out plain TF32, so every product is done as 3xTF32 (hi*hi + hi*lo + lo*hi,
//           fp32 accumulate), which is fp32-accurate to ~2^-22.
#include "common.cuh"
#include <stdlib.h>

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float apply_epilogue(float v, int c, int64_t row, int Cout, const insmos_epilogue_t& ep) {
    if (ep.scale) v = __fmaf_rn(v, __ldg(ep.scale + c), __ldg(ep.shift + c));
    if (ep.bias) v += __ldg(ep.bias + c);
    if (ep.residual) v += __ldg(ep.residual + row * Cout + c);
    if (ep.relu) v = fmaxf(v, 0.0f);
    return v;
}

// ------------------------------------------------------------------------------------------------
// algo 1: SIMT
#define SIMT_THREADS 256
__global__ void __launch_bounds__(SIMT_THREADS)
k_spconv_simt(const float* __restrict__ in, const float* __restrict__ W,
              const uint16_t* __restrict__ seg, const uint32_t* __restrict__ entries,
              float* __restrict__ out, int64_t n_out, int Cin, int Cout, int K, int TM, insmos_epilogue_t ep) {
    extern __shared__ float sm[];
    float* acc = sm;                                  // [TM*Cout]
    int* sseg = reinterpret_cast<int*>(acc + TM * Cout);   // [K+1]
    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const uint16_t* tseg = seg + tile * (K + 1);
    for (int k = tid; k <= K; k += SIMT_THREADS) sseg[k] = tseg[k];
    for (int i = tid; i < TM * Cout; i += SIMT_THREADS) acc[i] = 0.0f;
    __syncthreads();
    const uint32_t* tent = entries + tile * (int64_t)TM * K;
    for (int k = 0; k < K; ++k) {
        const int s0 = sseg[k], n = sseg[k + 1] - s0;
        if (n == 0) continue;                          // uniform across the block
        const float* Wk = W + (int64_t)k * Cin * Cout;
        for (int idx = tid; idx < n * Cout; idx += SIMT_THREADS) {
            const int p = idx / Cout, co = idx - p * Cout;
            const uint32_t e = __ldg(tent + s0 + p);
            const float* x = in + (int64_t)(e & INSMOS_ROW_MASK) * Cin;
            float a = 0.0f;
            for (int ci = 0; ci < Cin; ++ci) a = __fmaf_rn(__ldg(x + ci), __ldg(Wk + ci * Cout + co), a);
            acc[(e >> INSMOS_ROW_BITS) * Cout + co] += a;
        }
        __syncthreads();
    }
    const int64_t row0 = tile * TM;
    const int rows = (int)((n_out - row0) < TM ? (n_out - row0) : TM);
    for (int i = tid; i < rows * Cout; i += SIMT_THREADS) {
        const int r = i / Cout, c = i - r * Cout;
        out[(row0 + r) * Cout + c] = apply_epilogue(acc[i], c, row0 + r, Cout, ep);
    }
}

// ------------------------------------------------------------------------------------------------
// algo 2: tensor cores, 3xTF32, one warp per (tile, group of NT n-tiles)
//
// Weights are prepared once per layer (insmos_conv_prep_weights) into tensor-core FRAGMENT ORDER, already split
// into TF32 hi/lo:  wfrag[k][nt][ks][lane] = uint4{hi(b0), hi(b1), lo(b0), lo(b1)} with, for lane = 4g+t,
//   b0 = W[k][8ks+2t][8nt+g],  b1 = W[k][8ks+2t+1][8nt+g]   (zero beyond Cin / Cout)
// so the B operand of one mma is a single coalesced 16-byte load (L1/L2 resident), no conversions in the loop.
// The contraction index is permuted (k-slots t, t+4 <-> channels 2t, 2t+1) so that the A operand of a gathered
// row is one 8-byte load per 8 channels.  Every warp is independent (own tile rows x own channel columns of the
// shared-memory accumulator): no block barriers; thousands of warps per layer hide the gather latency, and the
// next chunk's rule-book entries are prefetched while the current chunk is multiplied.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float r = x - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void k_prep_weights(const float* __restrict__ W, int K, int Cin, int Cout, int KS, int NT8, uint4* __restrict__ wf) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)K * NT8 * KS * 32;
    if (idx >= total) return;
    const int lane = (int)(idx & 31);
    int64_t r = idx >> 5;
    const int ks = (int)(r % KS); r /= KS;
    const int nt = (int)(r % NT8);
    const int k = (int)(r / NT8);
    const int g = lane >> 2, t = lane & 3;
    const int c0 = ks * 8 + 2 * t, n = nt * 8 + g;
    float b0 = 0.f, b1 = 0.f;
    if (n < Cout) {
        if (c0 < Cin) b0 = W[((int64_t)k * Cin + c0) * Cout + n];
        if (c0 + 1 < Cin) b1 = W[((int64_t)k * Cin + c0 + 1) * Cout + n];
    }
    uint4 v;
    split_tf32(b0, v.x, v.z);
    split_tf32(b1, v.y, v.w);
    wf[idx] = v;
}

extern "C" int64_t insmos_conv_wfrag_elems(int32_t K, int32_t Cin, int32_t Cout) {
    return (int64_t)K * ((Cout + 7) / 8) * ((Cin + 7) / 8) * 32 * 4;      // number of 32-bit words
}
extern "C" int insmos_conv_prep_weights(const float* weight, int32_t K, int32_t Cin, int32_t Cout, void* wfrag, void* stream) {
    if (!weight || !wfrag || K <= 0 || Cin <= 0 || Cout <= 0) return INSMOS_ERR_INVALID_ARG;
    const int KS = (Cin + 7) / 8, NT8 = (Cout + 7) / 8;
    const int64_t total = (int64_t)K * NT8 * KS * 32;
    k_prep_weights<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(weight, K, Cin, Cout, KS, NT8, (uint4*)wfrag);
    INSMOS_CHECK_LAUNCH("k_prep_weights");
    return INSMOS_OK;
}

extern "C" int insmos_sparse_conv_fwd(const float* in, int64_t n_in, int32_t Cin,
                                      const float* weight, int32_t K, int32_t Cout,
                                      const uint16_t* seg, const uint32_t* entries, int32_t TM,
                                      float* out, int64_t n_out,
                                      const insmos_epilogue_t* ep_in, int32_t algo, void* stream) {
    (void)algo;                                              // SIMT fp32 reference path (see insmos_sparse_conv_fwd_tc)
    if ((n_in > 0 && !in) || !weight || !seg || !entries || (n_out > 0 && !out) || Cin <= 0 || Cout <= 0 || K <= 0 || n_out < 0 || n_in < 0)
        return INSMOS_ERR_INVALID_ARG;
    if (TM != 16 && TM != 32 && TM != 64 && TM != 128) return INSMOS_ERR_INVALID_ARG;
    if (n_in > (int64_t)INSMOS_ROW_MASK + 1) return INSMOS_ERR_UNSUPPORTED;
    insmos_epilogue_t ep = {nullptr, nullptr, nullptr, nullptr, 0};
    if (ep_in) ep = *ep_in;
    if (ep.scale && !ep.shift) return INSMOS_ERR_INVALID_ARG;
    if (n_out == 0) return INSMOS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = sizeof(float) * (size_t)TM * Cout + sizeof(int) * (size_t)(K + 1);
    if (smem > 220 * 1024) return INSMOS_ERR_UNSUPPORTED;
    static thread_local insmos_smem_cfg_t configured;
    INSMOS_CHECK_CUDA(insmos_ensure_smem(k_spconv_simt, smem, configured));
    k_spconv_simt<<<(unsigned)ceil_div64(n_out, TM), SIMT_THREADS, smem, st>>>(in, weight, seg, entries, out, n_out,
                                                                               Cin, Cout, K, TM, ep);
    INSMOS_CHECK_LAUNCH("k_spconv_simt");
    return INSMOS_OK;
}

// ------------------------------------------------------------------------------------------------
__global__ void k_linear(const float* __restrict__ in, const float* __restrict__ W, float* __restrict__ out,
                         int64_t n, int Cin, int Cout, insmos_epilogue_t ep) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * Cout) return;
    const int64_t row = idx / Cout; const int co = (int)(idx - row * Cout);
    const float* x = in + row * Cin;
    float a = 0.0f;
    for (int ci = 0; ci < Cin; ++ci) a = __fmaf_rn(__ldg(x + ci), __ldg(W + ci * Cout + co), a);
    out[idx] = apply_epilogue(a, co, row, Cout, ep);
}
// 8 lanes per row: lane l takes channels l, l+8, ... (a row's loads are coalesced 32-byte pieces), keeps COUTP partial
// sums in registers, the 8 partials are added with 3 shuffle steps.  The [Cin x Cout] weight sits in shared memory
// (row pitch COUTP+1: conflict free).  Replaces the thread-per-output kernel for Cout <= 32 (BEV heads 256 -> 11 on
// 75 k pixels: 125 -> ~25 us; the 1x1 shortcut convolutions of the 4D blocks).
template <int COUTP>
__global__ void __launch_bounds__(256)
k_linear_g8(const float* __restrict__ in, const float* __restrict__ W, float* __restrict__ out,
            int64_t n, int Cin, int Cout, insmos_epilogue_t ep) {
    extern __shared__ float sw[];                                  // [Cin][COUTP+1]
    for (int i = threadIdx.x; i < Cin * COUTP; i += blockDim.x) {
        const int ci = i / COUTP, co = i - ci * COUTP;
        sw[ci * (COUTP + 1) + co] = co < Cout ? W[ci * Cout + co] : 0.0f;
    }
    __syncthreads();
    const int l = threadIdx.x & 7;
    const int64_t rows_per_block = blockDim.x >> 3;
    // whole warps iterate together (the shuffles need all 32 lanes); rows past n are computed on zeros and not stored
    for (int64_t base = (int64_t)blockIdx.x * rows_per_block + ((threadIdx.x >> 5) << 2); base < n; base += (int64_t)gridDim.x * rows_per_block) {
        const int64_t row = base + ((threadIdx.x & 31) >> 3);
        const bool valid = row < n;
        const float* x = in + (valid ? row : 0) * Cin;
        float acc[COUTP];
#pragma unroll
        for (int c = 0; c < COUTP; ++c) acc[c] = 0.0f;
        // the residual of the channels this lane will store is fetched together with the row (one DRAM round trip, not two)
        float resv[COUTP / 8];
#pragma unroll
        for (int j = 0; j < COUTP / 8; ++j) {
            const int c = l + 8 * j;
            resv[j] = (valid && ep.residual && c < Cout) ? __ldg(ep.residual + row * Cout + c) : 0.0f;
        }
        int ci = l;
        for (; ci + 24 < Cin; ci += 32) {                            // 4 independent row loads in flight per lane
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = valid ? __ldg(x + ci + 8 * u) : 0.0f;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float* w = sw + (ci + 8 * u) * (COUTP + 1);
#pragma unroll
                for (int c = 0; c < COUTP; ++c) acc[c] = __fmaf_rn(v[u], w[c], acc[c]);
            }
        }
        for (; ci < Cin; ci += 8) {
            const float v = valid ? __ldg(x + ci) : 0.0f;
            const float* w = sw + ci * (COUTP + 1);
#pragma unroll
            for (int c = 0; c < COUTP; ++c) acc[c] = __fmaf_rn(v, w[c], acc[c]);
        }
#pragma unroll
        for (int c = 0; c < COUTP; ++c) {
            float v = acc[c];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            acc[c] = v;
        }
#pragma unroll
        for (int c = 0; c < COUTP; ++c)
            if (valid && (c & 7) == l && c < Cout) {
                float v = acc[c];
                if (ep.scale) v = __fmaf_rn(v, __ldg(ep.scale + c), __ldg(ep.shift + c));
                if (ep.bias) v += __ldg(ep.bias + c);
                if (ep.residual) v += resv[c >> 3];
                if (ep.relu) v = fmaxf(v, 0.0f);
                out[row * Cout + c] = v;
            }
    }
}

template <int COUTP>
static int launch_linear_g8(const float* in, const float* weight, float* out, int64_t n, int Cin, int Cout,
                            const insmos_epilogue_t& ep, cudaStream_t st) {
    const size_t smem = sizeof(float) * (size_t)Cin * (COUTP + 1);
    static thread_local insmos_smem_cfg_t configured;
    INSMOS_CHECK_CUDA(insmos_ensure_smem(k_linear_g8<COUTP>, smem, configured));
    int64_t blocks = ceil_div64(n, 32);
    // a large weight is staged once per block and the block strides over rows; a small one (<= 4 KB) is cheap to stage,
    // so every warp gets ONE group of 4 rows: these layers are latency bound (3 dependent memory round trips per row
    // group), thousands of independent warps hide it (377 k rows, 16 -> 8: 44 us with the capped grid)
    if ((size_t)Cin * COUTP > 1024 && blocks > 148 * 16) blocks = 148 * 16;
    k_linear_g8<COUTP><<<(unsigned)blocks, 256, smem, st>>>(in, weight, out, n, Cin, Cout, ep);
    INSMOS_CHECK_LAUNCH("k_linear_g8");
    return INSMOS_OK;
}

// One thread per row (ncu of k_linear_g8: 70-88 % issue-slot utilisation, two instructions per MAC -- an LDS per FFMA --
// plus 3 shuffles per output): the row arrives as float4 loads, the weights are read as broadcast LDS.128 (every lane of
// a warp reads the same address), COUTP accumulators stay in registers, no shuffles.  Cin % 4 == 0.
// 1x1 convolution / Linear: thread t of a block owns the rows {t, t + 128, ...} -- R rows per thread, so that every weight
// vector read from shared memory feeds R rows.  tools/micro/ffma_rate.cu: a broadcast LDS.128 costs four shared-memory
// wavefronts, "2 LDS.128 + 8 FFMA" per lane tops out at 47 % of the FMA pipe; with R rows per thread the same loads feed
// 8 R FFMAs.  The rows of a thread are 128 apart, so the lanes of a warp still read 32 consecutive rows (full sectors).
// Outputs leave as float4, epilogue constants are read once per thread.
template <int COUTP, int R>
__global__ void __launch_bounds__(128)
k_linear_row(const float* __restrict__ in, const float* __restrict__ W, float* __restrict__ out,
             int64_t n, int Cin, int Cout, insmos_epilogue_t ep) {
    extern __shared__ __align__(16) float sw[];                   // [Cin][COUTP] | scale[COUTP] | shift[COUTP] | bias[COUTP]
    const int64_t first = ep.first_row ? (int64_t)__ldg(ep.first_row) : 0;          // rows below are not needed (dead-row hint)
    if ((int64_t)(blockIdx.x + 1) * (128 * R) <= first) return;
    float* sc = sw + (size_t)Cin * COUTP;
    for (int i = threadIdx.x; i < Cin * COUTP; i += blockDim.x) {
        const int ci = i / COUTP, co = i - ci * COUTP;
        sw[i] = co < Cout ? W[ci * Cout + co] : 0.0f;
    }
    for (int c = threadIdx.x; c < COUTP; c += blockDim.x) {
        sc[c] = (ep.scale && c < Cout) ? ep.scale[c] : 1.0f;
        sc[COUTP + c] = (ep.scale && c < Cout) ? ep.shift[c] : 0.0f;
        sc[2 * COUTP + c] = (ep.bias && c < Cout) ? ep.bias[c] : 0.0f;
    }
    __syncthreads();
    const int64_t row0 = (int64_t)blockIdx.x * (128 * R) + threadIdx.x;
    const float4* x[R];
    bool ok[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        ok[r] = row0 + 128 * r < n && row0 + 128 * r >= first;
        x[r] = reinterpret_cast<const float4*>(in + (ok[r] ? row0 + 128 * r : 0) * Cin);
    }
    float acc[R][COUTP];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < COUTP; ++c) acc[r][c] = 0.0f;
    for (int c4 = 0; c4 < Cin / 4; ++c4) {
        float4 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = ok[r] ? __ldg(x[r] + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float4* w = reinterpret_cast<const float4*>(sw + (c4 * 4 + u) * COUTP);
#pragma unroll
            for (int q = 0; q < COUTP / 4; ++q) {
                const float4 ww = w[q];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float xv = u == 0 ? v[r].x : u == 1 ? v[r].y : u == 2 ? v[r].z : v[r].w;
                    acc[r][4 * q + 0] = __fmaf_rn(xv, ww.x, acc[r][4 * q + 0]);
                    acc[r][4 * q + 1] = __fmaf_rn(xv, ww.y, acc[r][4 * q + 1]);
                    acc[r][4 * q + 2] = __fmaf_rn(xv, ww.z, acc[r][4 * q + 2]);
                    acc[r][4 * q + 3] = __fmaf_rn(xv, ww.w, acc[r][4 * q + 3]);
                }
            }
        }
    }
    const bool vec = (Cout & 3) == 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (!ok[r]) continue;
        const int64_t row = row0 + 128 * r;
#pragma unroll
        for (int c = 0; c < COUTP; ++c) {
            float t = acc[r][c];
            if (ep.scale) t = __fmaf_rn(t, sc[c], sc[COUTP + c]);
            if (ep.bias) t += sc[2 * COUTP + c];
            if (ep.residual && c < Cout) t += __ldg(ep.residual + row * Cout + c);
            if (ep.relu) t = fmaxf(t, 0.0f);
            acc[r][c] = t;
        }
        if (vec) {
#pragma unroll
            for (int q = 0; q < COUTP / 4; ++q)
                if (4 * q < Cout)
                    *reinterpret_cast<float4*>(out + row * Cout + 4 * q) = make_float4(acc[r][4 * q], acc[r][4 * q + 1], acc[r][4 * q + 2], acc[r][4 * q + 3]);
        } else {
#pragma unroll
            for (int c = 0; c < COUTP; ++c)
                if (c < Cout) out[row * Cout + c] = acc[r][c];
        }
    }
}

template <int COUTP, int R>
static int launch_linear_row(const float* in, const float* weight, float* out, int64_t n, int Cin, int Cout,
                             const insmos_epilogue_t& ep, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)Cin * COUTP + 3 * COUTP);
    static thread_local insmos_smem_cfg_t configured;
    INSMOS_CHECK_CUDA(insmos_ensure_smem(k_linear_row<COUTP, R>, smem, configured));
    k_linear_row<COUTP, R><<<(unsigned)ceil_div64(n, 128 * R), 128, smem, st>>>(in, weight, out, n, Cin, Cout, ep);
    INSMOS_CHECK_LAUNCH("k_linear_row");
    return INSMOS_OK;
}

extern "C" int insmos_linear_fwd(const float* in, int64_t n, int32_t Cin, const float* weight, int32_t Cout,
                                 float* out, const insmos_epilogue_t* ep_in, void* stream) {
    if (!in || !weight || !out || Cin <= 0 || Cout <= 0 || n < 0) return INSMOS_ERR_INVALID_ARG;
    insmos_epilogue_t ep = {nullptr, nullptr, nullptr, nullptr, 0};
    if (ep_in) ep = *ep_in;
    if (ep.scale && !ep.shift) return INSMOS_ERR_INVALID_ARG;
    if (n == 0) return INSMOS_OK;
    if (Cout <= 32 && (Cin & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (size_t)Cin * 32 * sizeof(float) <= 64 * 1024 &&
        getenv("INSMOS_LINEAR_G8") == nullptr) {
        // rows per thread: as many as the accumulators allow while the grid still covers the SMs a few times
        const bool small = n < 148ll * 128 * 8;
        if (Cout <= 4) return small ? launch_linear_row<4, 2>(in, weight, out, n, Cin, Cout, ep, (cudaStream_t)stream)
                                    : launch_linear_row<4, 4>(in, weight, out, n, Cin, Cout, ep, (cudaStream_t)stream);
        if (Cout <= 8) return small ? launch_linear_row<8, 2>(in, weight, out, n, Cin, Cout, ep, (cudaStream_t)stream)
                                    : launch_linear_row<8, 4>(in, weight, out, n, Cin, Cout, ep, (cudaStream_t)stream);
        if (Cout <= 16) return small ? launch_linear_row<16, 2>(in, weight, out, n, Cin, Cout, ep, (cudaStream_t)stream)
                                     : launch_linear_row<16, 4>(in, weight, out, n, Cin, Cout, ep, (cudaStream_t)stream);
        return launch_linear_row<32, 2>(in, weight, out, n, Cin, Cout, ep, (cudaStream_t)stream);
    }
    if (Cout <= 32 && (size_t)Cin * 33 * sizeof(float) <= 160 * 1024) {
        if (Cout <= 8) return launch_linear_g8<8>(in, weight, out, n, Cin, Cout, ep, (cudaStream_t)stream);
        if (Cout <= 16) return launch_linear_g8<16>(in, weight, out, n, Cin, Cout, ep, (cudaStream_t)stream);
        return launch_linear_g8<32>(in, weight, out, n, Cin, Cout, ep, (cudaStream_t)stream);
    }
    k_linear<<<(unsigned)ceil_div64(n * Cout, 256), 256, 0, (cudaStream_t)stream>>>(in, weight, out, n, Cin, Cout, ep);
    INSMOS_CHECK_LAUNCH("k_linear");
    return INSMOS_OK;
}

__global__ void k_affine_act(const float* __restrict__ x, float* __restrict__ out, int64_t n, int C, insmos_epilogue_t ep) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * C) return;
    const int64_t row = idx / C; const int c = (int)(idx - row * C);
    out[idx] = apply_epilogue(x[idx], c, row, C, ep);
}
extern "C" int insmos_affine_act(const float* x, int64_t n, int32_t C, float* out,
                                 const insmos_epilogue_t* ep_in, void* stream) {
    if (!x || !out || C <= 0 || n < 0 || !ep_in) return INSMOS_ERR_INVALID_ARG;
    if (ep_in->scale && !ep_in->shift) return INSMOS_ERR_INVALID_ARG;
    if (n == 0) return INSMOS_OK;
    k_affine_act<<<(unsigned)ceil_div64(n * C, 256), 256, 0, (cudaStream_t)stream>>>(x, out, n, C, *ep_in);
    INSMOS_CHECK_LAUNCH("k_affine_act");
    return INSMOS_OK;
}

template <typename IDX>
__global__ void k_concat2(const float* __restrict__ a, int C1, const float* __restrict__ b, int C2, int64_t n,
                          float* __restrict__ out) {
    // IDX = uint32_t when n*(C1+C2) < 2^32: the 64-bit division of the flat index cost more than the copy itself
    const IDX C = (IDX)(C1 + C2);
    const IDX idx = (IDX)blockIdx.x * blockDim.x + threadIdx.x;
    if ((int64_t)idx >= n * (int64_t)C) return;
    const IDX row = idx / C; const int c = (int)(idx - row * C);
    out[idx] = (c < C1) ? a[(size_t)row * C1 + c] : b[(size_t)row * C2 + (c - C1)];
}
// float4 version for channel counts that are multiples of 4 (all concatenations of the fused graph): a quarter of the
// threads and index arithmetic (the scalar kernel ran at ~0.7 TB/s effective: 71 us for the 377 k x 16 concat)
__global__ void k_concat2_v4(const float4* __restrict__ a, int Q1, const float4* __restrict__ b, int Q2, uint32_t total4,
                             float4* __restrict__ out) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total4) return;
    const uint32_t Q = (uint32_t)(Q1 + Q2);
    const uint32_t row = idx / Q;
    const int q = (int)(idx - row * Q);
    out[idx] = (q < Q1) ? __ldg(a + (size_t)row * Q1 + q) : __ldg(b + (size_t)row * Q2 + (q - Q1));
}
extern "C" int insmos_concat2(const float* a, int32_t C1, const float* b, int32_t C2, int64_t n,
                              float* out, void* stream) {
    if (!a || !b || !out || C1 <= 0 || C2 <= 0 || n < 0) return INSMOS_ERR_INVALID_ARG;
    if (n == 0) return INSMOS_OK;
    const int64_t total = n * (C1 + C2);
    const bool aligned = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if ((C1 & 3) == 0 && (C2 & 3) == 0 && aligned && total / 4 < (1ll << 32))
        k_concat2_v4<<<(unsigned)ceil_div64(total / 4, 256), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4*>(a), C1 / 4, reinterpret_cast<const float4*>(b), C2 / 4, (uint32_t)(total / 4),
            reinterpret_cast<float4*>(out));
    else if (total < (1ll << 31))
        k_concat2<uint32_t><<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(a, C1, b, C2, n, out);
    else
        k_concat2<unsigned long long><<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(a, C1, b, C2, n, out);
    INSMOS_CHECK_LAUNCH("k_concat2");
    return INSMOS_OK;
}

__global__ void k_pairsum_add(const float* __restrict__ a, const float* __restrict__ b, int64_t total, float* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const float2 v = reinterpret_cast<const float2*>(b)[idx];
    const float s = __fadd_rn(v.x, v.y);
    out[idx] = a ? __fadd_rn(a[idx], s) : s;
}
extern "C" int insmos_pairsum_add(const float* a, const float* b, int64_t n, int32_t C, float* out, void* stream) {
    if (!b || !out || C <= 0 || n < 0) return INSMOS_ERR_INVALID_ARG;
    if (n == 0) return INSMOS_OK;
    k_pairsum_add<<<(unsigned)ceil_div64(n * C, 256), 256, 0, (cudaStream_t)stream>>>(a, b, n * C, out);
    INSMOS_CHECK_LAUNCH("k_pairsum_add");
    return INSMOS_OK;
}

__global__ void k_gather_rows(const float* __restrict__ src, int C, const int32_t* __restrict__ idx, int64_t n,
                              float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * C) return;
    const int64_t row = i / C; const int c = (int)(i - row * C);
    const int32_t s = idx[row];
    out[i] = (s >= 0) ? src[(int64_t)s * C + c] : 0.0f;
}
extern "C" int insmos_gather_rows(const float* src, int32_t C, const int32_t* idx, int64_t n, float* out, void* stream) {
    if (!src || !idx || !out || C <= 0 || n < 0) return INSMOS_ERR_INVALID_ARG;
    if (n == 0) return INSMOS_OK;
    k_gather_rows<<<(unsigned)ceil_div64(n * C, 256), 256, 0, (cudaStream_t)stream>>>(src, C, idx, n, out);
    INSMOS_CHECK_LAUNCH("k_gather_rows");
    return INSMOS_OK;
}

__global__ void k_seg_accum(const float* __restrict__ feat, int C, const int32_t* __restrict__ inverse, int64_t n,
                            float* out, int32_t* cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * C) return;
    const int64_t row = i / C; const int c = (int)(i - row * C);
    const int32_t r = inverse[row];
    if (r < 0) return;
    atomicAdd(out + (int64_t)r * C + c, feat[i]);
    if (c == 0) atomicAdd(cnt + r, 1);
}
__global__ void k_seg_div(float* out, const int32_t* __restrict__ cnt, int64_t n_rows, int C) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * C) return;
    const int32_t k = cnt[i / C];
    out[i] = __fdiv_rn(out[i], (float)(k < 1 ? 1 : k));
}
extern "C" int insmos_segment_mean(const float* feat, int32_t C, const int32_t* inverse, int64_t n,
                                   float* out, int32_t* cnt, int64_t n_rows, void* stream) {
    if (!feat || !inverse || !out || !cnt || C <= 0 || n < 0 || n_rows < 0) return INSMOS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rows == 0) return INSMOS_OK;
    INSMOS_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * n_rows * C, st));
    INSMOS_CHECK_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * n_rows, st));
    if (n > 0) {
        k_seg_accum<<<(unsigned)ceil_div64(n * C, 256), 256, 0, st>>>(feat, C, inverse, n, out, cnt);
        INSMOS_CHECK_LAUNCH("k_seg_accum");
    }
    k_seg_div<<<(unsigned)ceil_div64(n_rows * C, 256), 256, 0, st>>>(out, cnt, n_rows, C);
    INSMOS_CHECK_LAUNCH("k_seg_div");
    return INSMOS_OK;
}

__global__ void k_build_current(const float* __restrict__ points, int stride, const int32_t* __restrict__ cur_index,
                                int64_t n_cur, const int32_t* __restrict__ inverse, const float* __restrict__ vox_feat,
                                int Cfeat, int Cm, float* __restrict__ out) {
    const int CO = 4 + Cm;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cur * CO) return;
    const int64_t j = i / CO; const int c = (int)(i - j * CO);
    const int32_t p = cur_index[j];
    float v;
    if (c < 4) v = points[(int64_t)p * stride + c];
    else { const int32_t r = inverse[p]; v = (r >= 0) ? vox_feat[(int64_t)r * Cfeat + (c - 4)] : 0.0f; }
    out[i] = v;
}
extern "C" int insmos_build_current_points(const float* points, int32_t point_stride, const int32_t* cur_index,
                                           int64_t n_cur, const int32_t* inverse, const float* vox_feat, int32_t Cfeat,
                                           int32_t Cm, float* out, void* stream) {
    if (!points || !cur_index || !inverse || !vox_feat || !out || Cm <= 0 || Cm > Cfeat || point_stride < 4 || n_cur < 0)
        return INSMOS_ERR_INVALID_ARG;
    if (n_cur == 0) return INSMOS_OK;
    k_build_current<<<(unsigned)ceil_div64(n_cur * (4 + Cm), 256), 256, 0, (cudaStream_t)stream>>>(
        points, point_stride, cur_index, n_cur, inverse, vox_feat, Cfeat, Cm, out);
    INSMOS_CHECK_LAUNCH("k_build_current");
    return INSMOS_OK;
}

__global__ void k_dense_scatter(const float* __restrict__ feat, const int32_t* __restrict__ coords, int64_t n, int C,
                                int D, int H, int W, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * C) return;
    const int64_t row = i / C; const int c = (int)(i - row * C);
    const int32_t* p = coords + row * 4;
    const int z = p[1], y = p[2], x = p[3];
    if ((unsigned)z >= (unsigned)D || (unsigned)y >= (unsigned)H || (unsigned)x >= (unsigned)W) return;
    out[(((int64_t)c * D + z) * H + y) * W + x] = feat[i];
}
extern "C" int insmos_dense_scatter(const float* feat, const int32_t* coords, int64_t n, int32_t C,
                                    int32_t D, int32_t H, int32_t W, float* out, void* stream) {
    if (!feat || !coords || !out || C <= 0 || D <= 0 || H <= 0 || W <= 0 || n < 0) return INSMOS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    INSMOS_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)C * D * H * W, st));
    if (n == 0) return INSMOS_OK;
    k_dense_scatter<<<(unsigned)ceil_div64(n * C, 256), 256, 0, st>>>(feat, coords, n, C, D, H, W, out);
    INSMOS_CHECK_LAUNCH("k_dense_scatter");
    return INSMOS_OK;
}
