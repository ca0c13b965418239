// Sparse convolution, exact-fp32 path for the NARROW layers (MotionNet 4D U-Net and the 16-channel levels of the 3D U-Net):
// gather -> FFMA -> shared-memory accumulate, block-cooperative and phase-synchronous.   SURVEY.md section 8 rows a3 / a8 / a12.
//
// Why FFMA and not tensor cores here.  fp32 parity (logits within 1e-3 over ~60 layers) needs fp32-accurate products.  On
// sm_100a the legacy mma.sync TF32 path issues one HMMA.1688 per ~17 cycles per SM sub-partition (profiles/r01_conv_v2_sass_notes.md):
// with the 3xTF32 split that is ~45 TFLOP/s effective, BELOW the 72 TFLOP/s of plain FFMA, and tcgen05.mma has an
// ~80-cycle floor per instruction for N <= 128 (profiles/r01_umma_notes.md), which at N = Cout <= 32 is 4x waste before the
// 3x split.  These layers (K = 27/81/125 offsets, 8..48 -> 8..32 channels) are therefore computed exactly, on the FMA pipe.
//
// Decomposition.  A block owns a SUPER-TILE of G consecutive rule-book tiles (R = G*TM output rows, all Cout channels):
//   * acc[R][Cout] lives in shared memory for the whole kernel; every output row is written to HBM exactly once, with
//     BatchNorm / bias / residual / ReLU applied (no scatter, no atomics);
//   * the kernel offsets k are processed as PHASES, one __syncthreads per non-empty phase: W[k] (Cin x Cout, contiguous
//     in the [K,Cin,Cout] weight) is copied ONCE per block with cp.async into a double-buffered shared tile one phase
//     ahead, and read by every lane as broadcast LDS.128 (all lanes of a cout group read the same 16 bytes: one wavefront);
//   * the pairs of the phase's bucket (concatenated over the G tiles) are cut into chunks of PW pairs; a chunk is one
//     warp's work item: lane = (pair slot p < PW, cout group q < 32/PW); the lane gathers ITS pair's input row straight
//     from global memory into registers (Cin/4 LDG.128, rows are >= 32 contiguous bytes), keeps CL = Cout/(32/PW)
//     accumulators in registers seeded from acc[out_row][q*CL..], runs Cin x CL FFMAs and stores them back.  Inside a
//     bucket every output row occurs once, so the read-modify-write needs no atomics; the phase barrier orders buckets.
//   * each warp walks its own item list (k, chunk) with a register pipeline: rule-book entry two items ahead, gathered
//     row slice one step ahead -- the two dependent global-load latencies never sit on the critical path.
// Small chunks (PW = 8/16) raise the fill of the last chunk of a bucket (C2 cloud, R = 256: 0.79 at 32 pairs, 0.89 at 16)
// and give every warp of the block an item in most phases.
// The summation order is fixed (k ascending, channels ascending): results do not depend on scheduling.
#include "common.cuh"
#include <stdlib.h>

#define FMA_INVALID 0xffffffffu

struct FmaArgs {
    const float* in; const float* w; const uint16_t* seg; const uint32_t* entries; float* out;
    int64_t n_out; int n_tiles, Cin, K, TM, G;
    insmos_epilogue_t ep;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// acc[0..3] += x * w4  -- plain FFMA, or the packed fma.rn.f32x2 of sm_100 (FFMA2: two fp32 FMAs per lane and instruction)
template <bool F2>
__device__ __forceinline__ void fma4(float (&a)[4], float x, const float4& w) {
    if (F2) {
        unsigned long long xv, a0, a1, w0, w1;
        asm("mov.b64 %0, {%1,%1};" : "=l"(xv) : "f"(x));
        asm("mov.b64 %0, {%1,%2};" : "=l"(a0) : "f"(a[0]), "f"(a[1]));
        asm("mov.b64 %0, {%1,%2};" : "=l"(a1) : "f"(a[2]), "f"(a[3]));
        asm("mov.b64 %0, {%1,%2};" : "=l"(w0) : "f"(w.x), "f"(w.y));
        asm("mov.b64 %0, {%1,%2};" : "=l"(w1) : "f"(w.z), "f"(w.w));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a0) : "l"(xv), "l"(w0));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a1) : "l"(xv), "l"(w1));
        asm("mov.b64 {%0,%1}, %2;" : "=f"(a[0]), "=f"(a[1]) : "l"(a0));
        asm("mov.b64 {%0,%1}, %2;" : "=f"(a[2]), "=f"(a[3]) : "l"(a1));
    } else {
        a[0] = __fmaf_rn(x, w.x, a[0]); a[1] = __fmaf_rn(x, w.y, a[1]);
        a[2] = __fmaf_rn(x, w.z, a[2]); a[3] = __fmaf_rn(x, w.w, a[3]);
    }
}

// CS: input channels per register slice (Cin % CS == 0); PW: pairs per chunk; CL: output channels per lane.
template <int CS, int PW, int CL, bool F2>
__global__ void __launch_bounds__(256)
k_spconv_fma(FmaArgs p) {
    constexpr int NQ = 32 / PW;                              // cout groups per warp
    constexpr int COUT = NQ * CL;
    constexpr int AS = COUT + 4;                             // accumulator row stride in floats (+16 B: spreads the bank groups)
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, nwarps = nthreads >> 5;
    const int K = p.K, TM = p.TM, G = p.G, Cin = p.Cin;
    const int R = G * TM, NS = Cin / CS, WN = Cin * COUT;
    float* acc = sm;                                                         // [R][AS]
    float* wbuf = acc + (size_t)R * AS;                                      // [2][Cin][COUT]
    uint16_t* segs = reinterpret_cast<uint16_t*>(wbuf + 2 * (size_t)WN);     // [G][K+1]
    uint16_t* ntot = segs + G * (K + 1);                                     // [K+1] pairs of bucket k over the G tiles (ntot[K] = 0)
    const int tile0 = blockIdx.x * G;
    const int ntile = min(G, p.n_tiles - tile0);
    const int ps = lane % PW, q = lane / PW;                                 // pair slot, cout group

    for (int i = tid; i < G * (K + 1); i += nthreads) {
        const int g = i / (K + 1);
        segs[i] = (g < ntile) ? p.seg[(size_t)(tile0 + g) * (K + 1) + (i - g * (K + 1))] : (uint16_t)0;
    }
    {
        float4* a4 = reinterpret_cast<float4*>(acc);
        for (int i = tid; i < R * AS / 4; i += nthreads) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    for (int k = tid; k <= K; k += nthreads) {
        int n = 0;
        if (k < K) for (int g = 0; g < G; ++g) n += (int)segs[g * (K + 1) + k + 1] - (int)segs[g * (K + 1) + k];
        ntot[k] = (uint16_t)n;
    }
    __syncthreads();

    auto next_phase = [&](int k) { ++k; while (k < K && ntot[k] == 0) ++k; return k; };
    auto stage_w = [&](int k, int buf) {                                     // W[k] -> wbuf[buf], 16 B per cp.async
        const float4* src = reinterpret_cast<const float4*>(p.w + (size_t)k * WN);
        float4* dst = reinterpret_cast<float4*>(wbuf + (size_t)buf * WN);
        for (int i = tid; i < WN / 4; i += nthreads) cp_async16(dst + i, src + i);
        cp_async_commit();
    };
    // this warp's items: chunk c = warp, warp + nwarps, ... of every non-empty phase, in phase order
    auto advance = [&](int& k, int& c) {
        c += nwarps;
        while (k < K && c * PW >= (int)ntot[k]) { ++k; c = warp; }
    };
    const float* __restrict__ in = p.in;
    auto fetch_entry = [&](int k, int c, uint32_t& e, int& gbase) {
        e = FMA_INVALID; gbase = 0;
        if (k < K) {
            int pi = c * PW + ps;
            if (pi < (int)ntot[k]) {
                for (int g = 0; g < G; ++g) {
                    const int s0 = segs[g * (K + 1) + k], cnt = (int)segs[g * (K + 1) + k + 1] - s0;
                    if (pi < cnt) { e = __ldg(p.entries + (size_t)(tile0 + g) * TM * K + s0 + pi); gbase = g * TM; break; }
                    pi -= cnt;
                }
            }
        }
    };
    float xn[CS];                                                            // the NEXT row slice to be consumed
    auto load_slice = [&](uint32_t e, int s) {
        if (e != FMA_INVALID) {
            const float4* src = reinterpret_cast<const float4*>(in + (size_t)(e & INSMOS_ROW_MASK) * Cin + s * CS);
#pragma unroll
            for (int j = 0; j < CS / 4; ++j) {
                const float4 v = __ldg(src + j);
                xn[4 * j] = v.x; xn[4 * j + 1] = v.y; xn[4 * j + 2] = v.z; xn[4 * j + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < CS; ++j) xn[j] = 0.f;
        }
    };

    int k0 = -1, c0 = warp - nwarps, k1, c1, k2, c2, gb0, gb1, gb2;
    uint32_t e0, e1, e2;
    k0 = 0; while (k0 < K && ntot[k0] == 0) ++k0;                            // first non-empty phase
    const int kfirst = k0;
    c0 = warp;
    if (k0 < K && c0 * PW >= (int)ntot[k0]) advance(k0, c0);
    fetch_entry(k0, c0, e0, gb0);
    k1 = k0; c1 = c0; if (k1 < K) advance(k1, c1);
    fetch_entry(k1, c1, e1, gb1);
    k2 = k1; c2 = c1; if (k2 < K) advance(k2, c2);
    fetch_entry(k2, c2, e2, gb2);
    load_slice(e0, 0);

    int buf = 0;
    if (kfirst < K) stage_w(kfirst, 0);
    for (int k = kfirst; k < K; k = next_phase(k)) {
        cp_async_wait_all();
        __syncthreads();                                     // W[k] landed for everyone; all accumulates of the previous phase done
        const int kn = next_phase(k);
        if (kn < K) stage_w(kn, buf ^ 1);                    // its previous reader (the phase before k) has passed the barrier
        const float* wk = wbuf + (size_t)buf * WN + q * CL;
        while (k0 == k) {
            const bool valid = e0 != FMA_INVALID;
            float* arow = acc + (size_t)(gb0 + (int)(e0 >> INSMOS_ROW_BITS)) * AS + q * CL;
            float a[CL / 4][4];
#pragma unroll
            for (int j = 0; j < CL / 4; ++j) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) v = *reinterpret_cast<const float4*>(arow + 4 * j);
                a[j][0] = v.x; a[j][1] = v.y; a[j][2] = v.z; a[j][3] = v.w;
            }
            for (int s = 0; s < NS; ++s) {
                float x[CS];
#pragma unroll
                for (int j = 0; j < CS; ++j) x[j] = xn[j];
                if (s + 1 < NS) load_slice(e0, s + 1); else load_slice(e1, 0);       // one step ahead
                const float* ws = wk + (size_t)s * CS * COUT;
#pragma unroll
                for (int ci = 0; ci < CS; ++ci) {
#pragma unroll
                    for (int j = 0; j < CL / 4; ++j) {
                        const float4 w4 = *reinterpret_cast<const float4*>(ws + ci * COUT + 4 * j);
                        fma4<F2>(a[j], x[ci], w4);
                    }
                }
            }
            if (valid) {
#pragma unroll
                for (int j = 0; j < CL / 4; ++j)
                    *reinterpret_cast<float4*>(arow + 4 * j) = make_float4(a[j][0], a[j][1], a[j][2], a[j][3]);
            }
            k0 = k1; c0 = c1; e0 = e1; gb0 = gb1;
            k1 = k2; c1 = c2; e1 = e2; gb1 = gb2;
            if (k2 < K) advance(k2, c2);
            fetch_entry(k2, c2, e2, gb2);                                            // entries two items ahead
        }
        buf ^= 1;
    }
    __syncthreads();
    // epilogue: every output row leaves the SM once
    const int64_t row0 = (int64_t)tile0 * TM;
    const int rows = (int)((p.n_out - row0) < R ? (p.n_out - row0) : R);
    constexpr int C4 = COUT / 4;
    for (int i = tid; i < rows * C4; i += nthreads) {
        const int r = i / C4, c = (i - r * C4) * 4;
        float4 v = *reinterpret_cast<const float4*>(acc + (size_t)r * AS + c);
        float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float t = o[j];
            if (p.ep.scale) t = __fmaf_rn(t, __ldg(p.ep.scale + c + j), __ldg(p.ep.shift + c + j));
            if (p.ep.bias) t += __ldg(p.ep.bias + c + j);
            if (p.ep.residual) t += __ldg(p.ep.residual + (row0 + r) * COUT + c + j);
            if (p.ep.relu) t = fmaxf(t, 0.0f);
            o[j] = t;
        }
        *reinterpret_cast<float4*>(p.out + (row0 + r) * COUT + c) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <int CS, int PW, int CL, bool F2>
static int launch_fma(FmaArgs a, cudaStream_t st) {
    constexpr int COUT = (32 / PW) * CL;
    // rows per block: large super-tiles fill the chunks (a bucket holds ~0.2 R pairs in the 4D maps) but the grid must still
    // cover the 148 SMs a few times over
    int R = env_int("INSMOS_FMA_R", 0);
    if (R <= 0) R = a.n_out >= 148ll * 2 * 256 ? 256 : a.n_out >= 148ll * 128 ? 128 : 64;
    if (R < a.TM) R = a.TM;
    int G = R / a.TM;
    auto smem_of = [&](int g) {
        return sizeof(float) * ((size_t)g * a.TM * (COUT + 4) + 2 * (size_t)a.Cin * COUT) + sizeof(uint16_t) * ((size_t)g * (a.K + 1) + a.K + 2);
    };
    while (G > 1 && smem_of(G) > 100 * 1024) G /= 2;                    // >= 2 blocks per SM
    const size_t smem = (smem_of(G) + 15) & ~(size_t)15;
    if (smem > 220 * 1024) return INSMOS_ERR_UNSUPPORTED;
    a.G = G;
    int nwarps = env_int("INSMOS_FMA_WARPS", 0);
    if (nwarps != 2 && nwarps != 4 && nwarps != 8) nwarps = 8;
    static thread_local insmos_smem_cfg_t configured;
    INSMOS_CHECK_CUDA(insmos_ensure_smem(k_spconv_fma<CS, PW, CL, F2>, smem, configured));
    k_spconv_fma<CS, PW, CL, F2><<<(unsigned)ceil_div64(a.n_tiles, G), nwarps * 32, smem, st>>>(a);
    INSMOS_CHECK_LAUNCH("k_spconv_fma");
    return INSMOS_OK;
}

template <int PW, int CL>
static int dispatch_fma(const FmaArgs& a, cudaStream_t st) {
    const bool f2 = env_int("INSMOS_FMA_F2", 0) != 0;                  // packed FFMA2 variant (A/B switch, read per call)
    if (a.Cin % 16 == 0) return f2 ? launch_fma<16, PW, CL, true>(a, st) : launch_fma<16, PW, CL, false>(a, st);
    if (a.Cin % 8 == 0) return f2 ? launch_fma<8, PW, CL, true>(a, st) : launch_fma<8, PW, CL, false>(a, st);
    return INSMOS_ERR_UNSUPPORTED;
}

extern "C" int insmos_sparse_conv_fma_supported(int32_t K, int32_t Cin, int32_t Cout) {
    return (Cin % 8 == 0 && Cin >= 8 && Cin <= 256 && (Cout == 8 || Cout == 16 || Cout == 32) && K >= 1 && K <= 4096) ? 1 : 0;
}

extern "C" int insmos_sparse_conv_fwd_fma(const float* in, int64_t n_in, int32_t Cin,
                                          const float* weight, int32_t K, int32_t Cout,
                                          const uint16_t* seg, const uint32_t* entries, int32_t TM,
                                          float* out, int64_t n_out,
                                          const insmos_epilogue_t* ep_in, void* stream) {
    if ((n_in > 0 && !in) || !weight || !seg || !entries || (n_out > 0 && !out) || Cin <= 0 || Cout <= 0 || K <= 0 || n_out < 0 || n_in < 0)
        return INSMOS_ERR_INVALID_ARG;
    if (TM != 16 && TM != 32 && TM != 64 && TM != 128) return INSMOS_ERR_INVALID_ARG;
    if (n_in > (int64_t)INSMOS_ROW_MASK) return INSMOS_ERR_UNSUPPORTED;      // entry 0xffffffff is the 'no pair' marker
    if (!insmos_sparse_conv_fma_supported(K, Cin, Cout)) return INSMOS_ERR_UNSUPPORTED;
    if ((int64_t)TM * K >= 65536) return INSMOS_ERR_INVALID_ARG;
    FmaArgs a;
    a.in = in; a.w = weight; a.seg = seg; a.entries = entries; a.out = out;
    a.n_out = n_out; a.n_tiles = (int)ceil_div64(n_out, TM); a.Cin = Cin; a.K = K; a.TM = TM; a.G = 1;
    a.ep = insmos_epilogue_t{nullptr, nullptr, nullptr, nullptr, 0};
    if (ep_in) a.ep = *ep_in;
    if (a.ep.scale && !a.ep.shift) return INSMOS_ERR_INVALID_ARG;
    if (n_out == 0) return INSMOS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // chunk shape: lane = (pair slot, cout group).  INSMOS_FMA_PW overrides the pairs per chunk for A/B measurements.
    const int pw = env_int("INSMOS_FMA_PW", 0);
    if (Cout == 8) {
        if (pw == 32) return dispatch_fma<32, 8>(a, st);
        return dispatch_fma<16, 4>(a, st);
    }
    if (Cout == 16) {
        if (pw == 32) return dispatch_fma<32, 16>(a, st);
        if (pw == 8) return dispatch_fma<8, 4>(a, st);
        return dispatch_fma<16, 8>(a, st);
    }
    if (pw == 16) return dispatch_fma<16, 16>(a, st);
    return dispatch_fma<8, 8>(a, st);
}
