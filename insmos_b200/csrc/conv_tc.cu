// Sparse convolution, tensor-core path for the NARROW layers (3xTF32 on mma.sync m16n8k8).
//
// Work decomposition: a unit = (tile of TM output rows, group of NT n-tiles); wpt warps share a unit (bucket k -> warp
// k mod wpt, private partial accumulator tiles in shared memory, summed in the epilogue); the pairs of one rule-book
// bucket are packed 16 at a time into the M dimension; weights pre-arranged in fragment order and pre-split into TF32
// hi/lo (insmos_conv_prep_weights in conv.cu).
//   k_spconv_tc4<NT,KSC>  default: per-warp chunk list + counted loop (Cin = 8*KSC compile-time)
//   k_spconv_tc3<NT,KSC>  previous generation with a data-dependent bucket iterator; still the path for channel counts
//                         that are not a multiple of 8 (KSC = 0: runtime k-step loop) and the A/B switch INSMOS_TC_V3
//   k_spconv_tc_big<NT>   block-cooperative variant for >= 64 output channels when the tcgen05 kernel is not eligible
// History of the measurements that shaped them: profiles/r01_conv_v2_sass_notes.md, profiles/r01_umma_notes.md section 4.
// v3 notes (kept because tc3 is still compiled):
//   * the TF32 hi/lo split of the gathered activations is hi = (x + 0x1000) & 0xffffe000, lo = x - hi (3 instructions; the
//     tensor core ignores the low 13 mantissa bits of lo) instead of cvt.rna.tf32 (emulated, ~4 instr + NaN path);
//   * channel counts that are multiples of 8 up to 48 are compile-time (KSC = Cin/8): no per-chunk 64-bit address
//     arithmetic, fully unrolled k-steps;
//   * the weight fragments of a bucket are loaded once per bucket and kept in registers across its chunks;
//   * the next chunk's rule-book entries AND gathered feature rows are prefetched, so the two dependent global-load
//     latencies overlap the current chunk's mma + shared-memory accumulate.
#include "common.cuh"
#include <stdlib.h>

#define TC_WARPS 4

__device__ __forceinline__ float tc_epilogue(float v, int c, int64_t row, int Cout, const insmos_epilogue_t& ep) {
    if (ep.scale) v = __fmaf_rn(v, __ldg(ep.scale + c), __ldg(ep.shift + c));
    if (ep.bias) v += __ldg(ep.bias + c);
    if (ep.residual) v += __ldg(ep.residual + row * Cout + c);
    if (ep.relu) v = fmaxf(v, 0.0f);
    return v;
}
__device__ __forceinline__ void split_trunc(float x, uint32_t& hi, uint32_t& lo) {
    // round-to-nearest TF32 by integer add + mask (unbiased; a truncating split accumulates a coherent bias through
    // the layers: measured 1e-4 drift of the head scores).  lo = x - hi is exact (|lo| <= 2^-11 |x|); the tensor core
    // ignores its low 13 mantissa bits, a 2^-22 relative effect.
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32x(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct TcArgs {
    const float* in; const uint4* wf; const uint16_t* seg; const uint32_t* entries; float* out;
    int64_t n_out, n_tiles; int groups, Cin, Cout, K, TM, KS, NT8;
    int wpt;                                                 // warps sharing one (tile, channel group): bucket k goes to warp k % wpt
    int terms;                                               // TF32 products per fp32 product: 3 (default, fp32-accurate), 2 or 1 (accuracy study only)
    insmos_epilogue_t ep;
};

// chunk iterator over the non-empty buckets of one tile
struct ChunkIt {
    const int* sseg; int K, k, s0, n, c0, kstep;
    __device__ __forceinline__ bool next() {
        c0 += 16;
        while (c0 >= n) {
            if ((k += kstep) >= K) return false;
            s0 = sseg[k]; n = sseg[k + 1] - s0; c0 = 0;
        }
        return true;
    }
};

template <int NT, int KSC>
__global__ void __launch_bounds__(256)
k_spconv_tc3(TcArgs p) {
    constexpr int CW = NT * 8;
    constexpr int KSR = KSC > 0 ? KSC : 1;
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    // a UNIT = (tile of TM output rows, group of NT n-tiles) is shared by wpt warps: warp `sub` of the unit takes the
    // buckets k = sub, sub + wpt, ... into a private copy of the accumulator tile; the copies are summed in the
    // epilogue.  Large tiles keep the 16-pair chunks full (a bucket of a TM-row tile holds ~0.2*TM pairs in the 4D
    // maps), several warps per tile keep the SMs occupied (ncu of v3: 24 % warps active with one warp per tile).
    const int wpt = p.wpt;
    const int unit_local = warp / wpt, sub = warp - unit_local * wpt;
    const int64_t unit = (int64_t)blockIdx.x * (nwarps / wpt) + unit_local;
    const int64_t tile = unit / p.groups;
    const int grp = (int)(unit - tile * p.groups);
    const bool active = tile < p.n_tiles;
    const int TM = p.TM, K = p.K;
    const int Cin = KSC > 0 ? KSC * 8 : p.Cin;
    const int KS = KSC > 0 ? KSC : p.KS;
    float* acc = sm + (size_t)warp * TM * CW;
    int* sseg = reinterpret_cast<int*>(sm + (size_t)nwarps * TM * CW) + warp * (K + 1);
    if (active) {
        const uint16_t* tseg = p.seg + tile * (K + 1);
        for (int k = lane; k <= K; k += 32) sseg[k] = tseg[k];
        for (int i = lane; i < TM * CW; i += 32) acc[i] = 0.0f;
    }
    __syncwarp();
    const uint32_t* tent = p.entries + tile * (int64_t)TM * K;
    const int nt0 = grp * NT;
    const float* __restrict__ in = p.in;

    // three-stage software pipeline over chunks: entries of chunk i+2 and gathered rows of chunk i+1 are in flight
    // while chunk i is multiplied, so neither global-load latency sits on the critical path.
    struct Ent { uint32_t lo, hi; int k; bool vlo, vhi, ok; };
    ChunkIt it{sseg, K, sub - wpt, 0, 0, 0, wpt};
    auto fetch = [&]() -> Ent {
        Ent e; e.lo = 0u; e.hi = 0u; e.vlo = false; e.vhi = false;
        e.ok = active && it.next(); e.k = it.k;
        if (e.ok) {
            e.vlo = (it.c0 + g) < it.n; e.vhi = (it.c0 + g + 8) < it.n;
            if (e.vlo) e.lo = __ldg(tent + it.s0 + it.c0 + g);
            if (e.vhi) e.hi = __ldg(tent + it.s0 + it.c0 + g + 8);
        }
        return e;
    };
    float2 rlo[KSR], rhi[KSR];                               // rows of the NEXT chunk (compile-time path)
    auto load_rows = [&](const Ent& e) {
        if (KSC > 0) {
            const float* xl = in + (size_t)(e.lo & INSMOS_ROW_MASK) * (KSC * 8) + 2 * t;
            const float* xh = in + (size_t)(e.hi & INSMOS_ROW_MASK) * (KSC * 8) + 2 * t;
#pragma unroll
            for (int ks = 0; ks < KSR; ++ks) {
                rlo[ks] = e.vlo ? __ldg(reinterpret_cast<const float2*>(xl + ks * 8)) : make_float2(0.f, 0.f);
                rhi[ks] = e.vhi ? __ldg(reinterpret_cast<const float2*>(xh + ks * 8)) : make_float2(0.f, 0.f);
            }
        }
    };
    Ent E0 = fetch();
    Ent E1 = E0.ok ? fetch() : E0;
    if (E0.ok) load_rows(E0);

    int kb = -1;                                             // bucket whose weight fragments are in registers
    uint4 bfrag[NT][KSR];
    while (E0.ok) {
        const int kc = E0.k;
        const bool cv_lo = E0.vlo, cv_hi = E0.vhi;
        const uint32_t ce_lo = E0.lo, ce_hi = E0.hi;
        float2 clo[KSR], chi[KSR];
#pragma unroll
        for (int ks = 0; ks < KSR; ++ks) { clo[ks] = rlo[ks]; chi[ks] = rhi[ks]; }
        Ent E2 = E1.ok ? fetch() : E1;                       // entries of chunk i+2
        if (E1.ok) load_rows(E1);                            // rows of chunk i+1 (its entries arrived last iteration)
        float d[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) { d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.0f; }

        if (KSC > 0) {
            if (kc != kb) {                                  // new bucket: fetch its weight fragments once
                const uint4* wk = p.wf + ((int64_t)kc * p.NT8 + nt0) * KSC * 32 + lane;
#pragma unroll
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int ks = 0; ks < KSR; ++ks) bfrag[j][ks] = __ldg(wk + (j * KSC + ks) * 32);
                kb = kc;
            }
#pragma unroll
            for (int ks = 0; ks < KSR; ++ks) {
                uint32_t ah[4], al[4];
                split_trunc(clo[ks].x, ah[0], al[0]); split_trunc(chi[ks].x, ah[1], al[1]);
                split_trunc(clo[ks].y, ah[2], al[2]); split_trunc(chi[ks].y, ah[3], al[3]);
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    mma_tf32x(d[j], al, bfrag[j][ks].x, bfrag[j][ks].y);
                    mma_tf32x(d[j], ah, bfrag[j][ks].z, bfrag[j][ks].w);
                    mma_tf32x(d[j], ah, bfrag[j][ks].x, bfrag[j][ks].y);
                }
            }
        } else {
            const bool even = (Cin & 1) == 0;
            const float* x_lo = in + (size_t)(ce_lo & INSMOS_ROW_MASK) * Cin;
            const float* x_hi = in + (size_t)(ce_hi & INSMOS_ROW_MASK) * Cin;
            const uint4* wk = p.wf + ((int64_t)kc * p.NT8 + nt0) * KS * 32 + lane;
            constexpr int KG = 8;                            // k-steps whose loads are all in flight together
            if (KS <= 3) {                                   // few k-steps: the plain loop (no group overhead)
                for (int ks = 0; ks < KS; ++ks) {
                    const int col = ks * 8 + 2 * t;
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                    if (col < Cin) { if (cv_lo) a0 = __ldg(x_lo + col); if (cv_hi) a1 = __ldg(x_hi + col); }
                    if (col + 1 < Cin) { if (cv_lo) a2 = __ldg(x_lo + col + 1); if (cv_hi) a3 = __ldg(x_hi + col + 1); }
                    uint32_t ah[4], al[4];
                    split_trunc(a0, ah[0], al[0]); split_trunc(a1, ah[1], al[1]);
                    split_trunc(a2, ah[2], al[2]); split_trunc(a3, ah[3], al[3]);
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const uint4 b = __ldg(wk + ((int64_t)j * KS + ks) * 32);
                        mma_tf32x(d[j], al, b.x, b.y);
                        mma_tf32x(d[j], ah, b.z, b.w);
                        mma_tf32x(d[j], ah, b.x, b.y);
                    }
                }
            } else
            for (int ks0 = 0; ks0 < KS; ks0 += KG) {
                float2 rl[KG], rh[KG];
                uint4 b[KG][NT];
#pragma unroll
                for (int u = 0; u < KG; ++u) {
                    const int col = (ks0 + u) * 8 + 2 * t;
                    rl[u] = make_float2(0.f, 0.f); rh[u] = make_float2(0.f, 0.f);
                    if (even) {
                        if (col < Cin) {
                            if (cv_lo) rl[u] = __ldg(reinterpret_cast<const float2*>(x_lo + col));
                            if (cv_hi) rh[u] = __ldg(reinterpret_cast<const float2*>(x_hi + col));
                        }
                    } else {
                        if (col < Cin) { if (cv_lo) rl[u].x = __ldg(x_lo + col); if (cv_hi) rh[u].x = __ldg(x_hi + col); }
                        if (col + 1 < Cin) { if (cv_lo) rl[u].y = __ldg(x_lo + col + 1); if (cv_hi) rh[u].y = __ldg(x_hi + col + 1); }
                    }
#pragma unroll
                    for (int j = 0; j < NT; ++j)
                        b[u][j] = (ks0 + u < KS) ? __ldg(wk + ((int64_t)j * KS + ks0 + u) * 32) : make_uint4(0u, 0u, 0u, 0u);
                }
#pragma unroll
                for (int u = 0; u < KG; ++u) {
                    if (ks0 + u < KS) {
                        uint32_t ah[4], al[4];
                        split_trunc(rl[u].x, ah[0], al[0]); split_trunc(rh[u].x, ah[1], al[1]);
                        split_trunc(rl[u].y, ah[2], al[2]); split_trunc(rh[u].y, ah[3], al[3]);
#pragma unroll
                        for (int j = 0; j < NT; ++j) {
                            mma_tf32x(d[j], al, b[u][j].x, b[u][j].y);
                            mma_tf32x(d[j], ah, b[u][j].z, b[u][j].w);
                            mma_tf32x(d[j], ah, b[u][j].x, b[u][j].y);
                        }
                    }
                }
            }
        }
        // accumulate into the tile (within a bucket every output row occurs once: plain read-modify-write)
        const int r_lo = (int)(ce_lo >> INSMOS_ROW_BITS) * CW + 2 * t;
        const int r_hi = (int)(ce_hi >> INSMOS_ROW_BITS) * CW + 2 * t;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            if (cv_lo) {
                float2* q = reinterpret_cast<float2*>(acc + r_lo + j * 8);
                float2 v = *q; v.x += d[j][0]; v.y += d[j][1]; *q = v;
            }
            if (cv_hi) {
                float2* q = reinterpret_cast<float2*>(acc + r_hi + j * 8);
                float2 v = *q; v.x += d[j][2]; v.y += d[j][3]; *q = v;
            }
        }
        __syncwarp();
        E0 = E1; E1 = E2;
    }
    __syncthreads();                                         // the unit's wpt partial tiles are complete
    if (!active) return;
    const int64_t row0 = tile * TM;
    const int rows = (int)((p.n_out - row0) < TM ? (p.n_out - row0) : TM);
    const int cbase = nt0 * 8;
    const float* acc0 = sm + (size_t)(unit_local * wpt) * TM * CW;
    for (int i = sub * 32 + lane; i < rows * CW; i += wpt * 32) {
        const int r = i / CW, c = cbase + (i % CW);
        float v = acc0[i];
        for (int w = 1; w < wpt; ++w) v += acc0[(size_t)w * TM * CW + i];
        if (c < p.Cout) p.out[(row0 + r) * p.Cout + c] = tc_epilogue(v, c, row0 + r, p.Cout, p.ep);
    }
}

// ------------------------------------------------------------------------------------------------
// v4: same decomposition and arithmetic as v3, without the data-dependent chunk iterator.
// ncu source page of v3 on the 8->8 K=81 layer (profiles/r01_conv_v4_notes.md): 169 warp instructions per 16-pair
// chunk, 43 % of them in the bucket iterator / pipeline-state rotation, 3 HMMA.  v4 first writes the warp's CHUNK
// LIST (one 32-bit descriptor k | start << 7 | count << 23 per 16-pair chunk of its buckets) into shared memory with
// a ballot-free prefix scan, then runs a plain counted loop over it: descriptor -> entries (prefetched two chunks
// ahead) -> gathered rows (one chunk ahead) -> split -> mma -> shared-memory accumulate.
#define TC4_INVALID 0xffffffffu
template <int NT, int KSC>
__global__ void __launch_bounds__(256, 2)                        // <= 128 registers: two 8-warp blocks per SM (ncu: <2,6> had 134 -> 12.5 % warps active)
k_spconv_tc4(TcArgs p) {
    constexpr int CW = NT * 8;
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wpt = p.wpt;
    const int unit_local = warp / wpt, sub = warp - unit_local * wpt;
    const int unit = (int)blockIdx.x * (nwarps / wpt) + unit_local;
    const int tile = unit / p.groups;
    const int grp = unit - tile * p.groups;
    const bool active = tile < (int)p.n_tiles;
    const int TM = p.TM, K = p.K;
    const int nbk = (K + wpt - 1) / wpt;                             // buckets of this warp: k = sub, sub + wpt, ...
    const int cap = nbk * (1 + TM / 16);                             // chunk-list capacity per warp
    float* acc = sm + (size_t)warp * TM * CW;
    uint32_t* list = reinterpret_cast<uint32_t*>(sm + (size_t)nwarps * TM * CW) + (size_t)warp * cap;
    int nch = 0;
    if (active) {
        for (int i = lane; i < TM * CW; i += 32) acc[i] = 0.0f;
        const uint16_t* tseg = p.seg + (size_t)tile * (K + 1);
        for (int j0 = 0; j0 < nbk; j0 += 32) {
            const int k = sub + wpt * (j0 + lane);
            int s0 = 0, n = 0;
            if (j0 + lane < nbk && k < K) { s0 = tseg[k]; n = (int)tseg[k + 1] - s0; }
            const int c = (n + 15) >> 4;
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
            int pos = nch + inc - c;
            for (int i = 0; i < c; ++i) {
                const int cnt = min(16, n - 16 * i);
                list[pos + i] = (uint32_t)k | ((uint32_t)(s0 + 16 * i) << 7) | ((uint32_t)cnt << 23);
            }
            nch += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    __syncwarp();
    const uint32_t* tent = p.entries + (size_t)tile * TM * K;
    const int nt0 = grp * NT;
    const float* __restrict__ in = p.in;
    auto ld_ent = [&](int ci, uint32_t& lo, uint32_t& hi, int& k) {
        lo = TC4_INVALID; hi = TC4_INVALID; k = -1;
        if (ci < nch) {
            const uint32_t d = list[ci];
            const int start = (int)((d >> 7) & 0xffffu), cnt = (int)(d >> 23);
            k = (int)(d & 127u);
            if (g < cnt) lo = __ldg(tent + start + g);
            if (g + 8 < cnt) hi = __ldg(tent + start + g + 8);
        }
    };
    float2 rlo[KSC], rhi[KSC];
    auto ld_rows = [&](uint32_t lo, uint32_t hi) {
        const float* xl = in + (size_t)(lo & INSMOS_ROW_MASK) * (KSC * 8) + 2 * t;
        const float* xh = in + (size_t)(hi & INSMOS_ROW_MASK) * (KSC * 8) + 2 * t;
#pragma unroll
        for (int ks = 0; ks < KSC; ++ks) {
            rlo[ks] = (lo != TC4_INVALID) ? __ldg(reinterpret_cast<const float2*>(xl + ks * 8)) : make_float2(0.f, 0.f);
            rhi[ks] = (hi != TC4_INVALID) ? __ldg(reinterpret_cast<const float2*>(xh + ks * 8)) : make_float2(0.f, 0.f);
        }
    };
    uint32_t e0lo, e0hi, e1lo, e1hi; int k0, k1;
    ld_ent(0, e0lo, e0hi, k0);
    ld_ent(1, e1lo, e1hi, k1);
    if (nch > 0) ld_rows(e0lo, e0hi);
    int kb = -1;
    uint4 bfrag[NT][KSC];
    for (int ci = 0; ci < nch; ++ci) {
        float2 clo[KSC], chi[KSC];
#pragma unroll
        for (int ks = 0; ks < KSC; ++ks) { clo[ks] = rlo[ks]; chi[ks] = rhi[ks]; }
        uint32_t e2lo, e2hi; int k2;
        ld_ent(ci + 2, e2lo, e2hi, k2);                              // entries two chunks ahead
        if (ci + 1 < nch) ld_rows(e1lo, e1hi);                       // rows one chunk ahead
        if (k0 != kb) {                                              // new bucket: its weight fragments
            const uint4* wk = p.wf + ((size_t)k0 * p.NT8 + nt0) * KSC * 32 + lane;
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int ks = 0; ks < KSC; ++ks) bfrag[j][ks] = __ldg(wk + (j * KSC + ks) * 32);
            kb = k0;
        }
        float d[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) { d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.0f; }
#pragma unroll
        for (int ks = 0; ks < KSC; ++ks) {
            uint32_t ah[4], al[4];
            split_trunc(clo[ks].x, ah[0], al[0]); split_trunc(chi[ks].x, ah[1], al[1]);
            split_trunc(clo[ks].y, ah[2], al[2]); split_trunc(chi[ks].y, ah[3], al[3]);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                if (p.terms >= 2) mma_tf32x(d[j], al, bfrag[j][ks].x, bfrag[j][ks].y);     // x_lo * w_hi
                if (p.terms >= 3) mma_tf32x(d[j], ah, bfrag[j][ks].z, bfrag[j][ks].w);     // x_hi * w_lo
                mma_tf32x(d[j], ah, bfrag[j][ks].x, bfrag[j][ks].y);                        // x_hi * w_hi
            }
        }
        // within a bucket every output row occurs once: plain read-modify-write of the warp's private tile
        if (e0lo != TC4_INVALID) {
            float* q0 = acc + (int)(e0lo >> INSMOS_ROW_BITS) * CW + 2 * t;
#pragma unroll
            for (int j = 0; j < NT; ++j) { float2* q = reinterpret_cast<float2*>(q0 + j * 8); float2 v = *q; v.x += d[j][0]; v.y += d[j][1]; *q = v; }
        }
        if (e0hi != TC4_INVALID) {
            float* q0 = acc + (int)(e0hi >> INSMOS_ROW_BITS) * CW + 2 * t;
#pragma unroll
            for (int j = 0; j < NT; ++j) { float2* q = reinterpret_cast<float2*>(q0 + j * 8); float2 v = *q; v.x += d[j][2]; v.y += d[j][3]; *q = v; }
        }
        __syncwarp();
        e0lo = e1lo; e0hi = e1hi; k0 = k1;
        e1lo = e2lo; e1hi = e2hi; k1 = k2;
    }
    __syncthreads();                                                 // the unit's wpt partial tiles are complete
    if (!active) return;
    const int64_t row0 = (int64_t)tile * TM;
    const int rows = (int)((p.n_out - row0) < TM ? (p.n_out - row0) : TM);
    const int cbase = nt0 * 8;
    const float* acc0 = sm + (size_t)(unit_local * wpt) * TM * CW;
    for (int i = sub * 32 + lane; i < rows * CW; i += wpt * 32) {
        const int r = i / CW, c = cbase + (i % CW);
        float v = acc0[i];
        for (int w = 1; w < wpt; ++w) v += acc0[(size_t)w * TM * CW + i];
        if (c < p.Cout) p.out[(row0 + r) * p.Cout + c] = tc_epilogue(v, c, row0 + r, p.Cout, p.ep);
    }
}

// ------------------------------------------------------------------------------------------------
// Large-channel layers (Cout >= 32): block-cooperative variant.
// With one warp per (tile, 8-channel group) the [Cin x 8] weight fragments of every bucket are re-fetched by
// every warp: for 128->128, K=27 on 6 k rows that is ~1.3 GB of L2->SM traffic per layer (measured 363 us).
// Here a block of 8 warps owns a SUPER-TILE (G consecutive rule-book tiles, ~128 rows) x a slice of NT n-tiles;
// per bucket the slice's weight fragments (contiguous in the fragment-ordered array) are staged ONCE in shared
// memory, the bucket's 16-pair chunks are dealt round-robin to the warps, each warp gathers its rows once and
// multiplies them against all NT n-tiles.  Chunks of one bucket touch distinct output rows, so the shared
// accumulator needs no atomics; a block barrier separates buckets.
#define BIG_WARPS 8
template <int NT>
__global__ void __launch_bounds__(BIG_WARPS * 32, 2)
k_spconv_tc_big(TcArgs p, int G, int n_slices) {
    constexpr int CN = NT * 8;
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int64_t stile = blockIdx.x / n_slices;
    const int slice = (int)(blockIdx.x - stile * n_slices);
    const int TM = p.TM, K = p.K, KS = p.KS, Cin = p.Cin;
    const int64_t tile0 = stile * G;
    const int ntile = (int)((p.n_tiles - tile0) < G ? (p.n_tiles - tile0) : G);
    const int nt0 = slice * NT;
    float* acc = sm;                                                    // [G*TM][CN]
    uint4* wbuf = reinterpret_cast<uint4*>(acc + (size_t)G * TM * CN);   // [NT][KS][32]
    int* ssegs = reinterpret_cast<int*>(wbuf + (size_t)NT * KS * 32);    // [G][K+1]
    for (int i = threadIdx.x; i < G * (K + 1); i += BIG_WARPS * 32) {
        const int gi = i / (K + 1), k = i - gi * (K + 1);
        ssegs[i] = (gi < ntile) ? p.seg[(tile0 + gi) * (K + 1) + k] : 0;
    }
    for (int i = threadIdx.x; i < G * TM * CN; i += BIG_WARPS * 32) acc[i] = 0.0f;
    __syncthreads();
    const bool even = (Cin & 1) == 0;
    const float* __restrict__ in = p.in;
    const int wn = NT * KS * 32;                                         // uint4 per weight slice
    for (int k = 0; k < K; ++k) {
        int tot = 0;
        for (int gi = 0; gi < ntile; ++gi) tot += ssegs[gi * (K + 1) + k + 1] - ssegs[gi * (K + 1) + k];
        if (tot == 0) continue;                                          // uniform across the block
        const uint4* wk = p.wf + ((int64_t)k * p.NT8 + nt0) * KS * 32;
        for (int i = threadIdx.x; i < wn; i += BIG_WARPS * 32) wbuf[i] = __ldg(wk + i);
        __syncthreads();
        int cid = 0;
        for (int gi = 0; gi < ntile; ++gi) {
            const int s0 = ssegs[gi * (K + 1) + k], n = ssegs[gi * (K + 1) + k + 1] - s0;
            const uint32_t* tent = p.entries + (tile0 + gi) * (int64_t)TM * K + s0;
            for (int c0 = 0; c0 < n; c0 += 16, ++cid) {
                if ((cid & (BIG_WARPS - 1)) != warp) continue;
                const bool v_lo = (c0 + g) < n, v_hi = (c0 + g + 8) < n;
                const uint32_t e_lo = v_lo ? __ldg(tent + c0 + g) : 0u;
                const uint32_t e_hi = v_hi ? __ldg(tent + c0 + g + 8) : 0u;
                const float* x_lo = in + (size_t)(e_lo & INSMOS_ROW_MASK) * Cin;
                const float* x_hi = in + (size_t)(e_hi & INSMOS_ROW_MASK) * Cin;
                float d[NT][4];
#pragma unroll
                for (int j = 0; j < NT; ++j) { d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.0f; }
                // k-steps in groups of KG: ALL gathered-row loads of a group are issued before the first is consumed,
                // so a chunk pays ~one L2 round trip per group instead of one per k-step
                constexpr int KG = 16;
                for (int ks0 = 0; ks0 < KS; ks0 += KG) {
                    float2 rl[KG], rh[KG];
#pragma unroll
                    for (int u = 0; u < KG; ++u) {
                        const int col = (ks0 + u) * 8 + 2 * t;
                        rl[u] = make_float2(0.f, 0.f); rh[u] = make_float2(0.f, 0.f);
                        if (even) {
                            if (col < Cin) {
                                if (v_lo) rl[u] = __ldg(reinterpret_cast<const float2*>(x_lo + col));
                                if (v_hi) rh[u] = __ldg(reinterpret_cast<const float2*>(x_hi + col));
                            }
                        } else {
                            if (col < Cin) { if (v_lo) rl[u].x = __ldg(x_lo + col); if (v_hi) rh[u].x = __ldg(x_hi + col); }
                            if (col + 1 < Cin) { if (v_lo) rl[u].y = __ldg(x_lo + col + 1); if (v_hi) rh[u].y = __ldg(x_hi + col + 1); }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < KG; ++u) {
                        const int ks = ks0 + u;
                        if (ks < KS) {
                            uint32_t ah[4], al[4];
                            split_trunc(rl[u].x, ah[0], al[0]); split_trunc(rh[u].x, ah[1], al[1]);
                            split_trunc(rl[u].y, ah[2], al[2]); split_trunc(rh[u].y, ah[3], al[3]);
#pragma unroll
                            for (int j = 0; j < NT; ++j) {
                                const uint4 b = wbuf[(j * KS + ks) * 32 + lane];
                                mma_tf32x(d[j], al, b.x, b.y);
                                mma_tf32x(d[j], ah, b.z, b.w);
                                mma_tf32x(d[j], ah, b.x, b.y);
                            }
                        }
                    }
                }
                const int r_lo = (gi * TM + (int)(e_lo >> INSMOS_ROW_BITS)) * CN + 2 * t;
                const int r_hi = (gi * TM + (int)(e_hi >> INSMOS_ROW_BITS)) * CN + 2 * t;
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    if (v_lo) { float2* q = reinterpret_cast<float2*>(acc + r_lo + j * 8); float2 v = *q; v.x += d[j][0]; v.y += d[j][1]; *q = v; }
                    if (v_hi) { float2* q = reinterpret_cast<float2*>(acc + r_hi + j * 8); float2 v = *q; v.x += d[j][2]; v.y += d[j][3]; *q = v; }
                }
            }
        }
        __syncthreads();                                                 // bucket done: acc rows + wbuf reusable
    }
    const int64_t row0 = tile0 * TM;
    const int64_t rows_left = p.n_out - row0;
    const int rows = (int)(rows_left < (int64_t)G * TM ? rows_left : (int64_t)G * TM);
    const int cbase = nt0 * 8;
    for (int i = threadIdx.x; i < rows * CN; i += BIG_WARPS * 32) {
        const int r = i / CN, c = cbase + (i % CN);
        if (c < p.Cout) p.out[(row0 + r) * p.Cout + c] = tc_epilogue(acc[i], c, row0 + r, p.Cout, p.ep);
    }
}

static int launch_tc_big(const TcArgs& a, cudaStream_t st) {
    constexpr int NT = 4;
    int G = 128 / a.TM; if (G < 1) G = 1;
    const int n_slices = (a.NT8 + NT - 1) / NT;
    const size_t smem = sizeof(float) * (size_t)G * a.TM * NT * 8 + sizeof(uint4) * (size_t)NT * a.KS * 32 +
                        sizeof(int) * (size_t)G * (a.K + 1);
    if (smem > 220 * 1024) return INSMOS_ERR_UNSUPPORTED;
    static thread_local insmos_smem_cfg_t configured;
    INSMOS_CHECK_CUDA(insmos_ensure_smem(k_spconv_tc_big<NT>, smem, configured));
    const int64_t stiles = ceil_div64(a.n_tiles, G);
    k_spconv_tc_big<NT><<<(unsigned)(stiles * n_slices), BIG_WARPS * 32, smem, st>>>(a, G, n_slices);
    INSMOS_CHECK_LAUNCH("k_spconv_tc_big");
    return INSMOS_OK;
}

template <int NT, int KSC>
static int launch_tc3(TcArgs a, cudaStream_t st) {
    // warps per unit: enough warps in flight to cover the gather latency on 148 SMs (~48 resident warps each)
    const int64_t units = a.n_tiles * a.groups;
    int wpt = 1;
    while (wpt < 8 && units * wpt < 8192 && wpt * 2 <= a.K) wpt *= 2;
    if (const char* e = getenv("INSMOS_WPT")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) wpt = v; }
    while (wpt > 1 && sizeof(float) * (size_t)(wpt > 4 ? wpt : 4) * a.TM * NT * 8 > 96 * 1024) wpt /= 2;
    a.wpt = wpt;
    const int nwarps = wpt > TC_WARPS ? wpt : TC_WARPS;
    const size_t smem = sizeof(float) * (size_t)nwarps * a.TM * NT * 8 + sizeof(int) * (size_t)nwarps * (a.K + 1);
    if (smem > 220 * 1024) return INSMOS_ERR_UNSUPPORTED;
    static thread_local insmos_smem_cfg_t configured;
    INSMOS_CHECK_CUDA(insmos_ensure_smem(k_spconv_tc3<NT, KSC>, smem, configured));
    k_spconv_tc3<NT, KSC><<<(unsigned)ceil_div64(units, nwarps / wpt), nwarps * 32, smem, st>>>(a);
    INSMOS_CHECK_LAUNCH("k_spconv_tc3");
    return INSMOS_OK;
}

static int choose_wpt(const TcArgs& a, int NT) {
    const int64_t units = a.n_tiles * a.groups;
    int wpt = 1;
    while (wpt < 8 && units * wpt < 8192 && wpt * 2 <= a.K) wpt *= 2;
    if (const char* e = getenv("INSMOS_WPT")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) wpt = v; }
    while (wpt > 1 && sizeof(float) * (size_t)(wpt > 4 ? wpt : 4) * a.TM * NT * 8 > 96 * 1024) wpt /= 2;
    return wpt;
}

template <int NT, int KSC>
static int launch_tc4(TcArgs a, cudaStream_t st) {
    static_assert(KSC > 0, "v4 needs a compile-time channel count");
    const int64_t units = a.n_tiles * a.groups;
    if (a.K > 127 || (int64_t)a.TM * a.K >= 65536 || units >= (1ll << 30)) return launch_tc3<NT, KSC>(a, st);
    const int wpt = choose_wpt(a, NT);
    a.wpt = wpt;
    const int nwarps = wpt > TC_WARPS ? wpt : TC_WARPS;
    const int nbk = (a.K + wpt - 1) / wpt;
    const size_t smem = sizeof(float) * (size_t)nwarps * a.TM * NT * 8 + sizeof(uint32_t) * (size_t)nwarps * nbk * (1 + a.TM / 16);
    if (smem > 220 * 1024) return launch_tc3<NT, KSC>(a, st);
    static thread_local insmos_smem_cfg_t configured;
    INSMOS_CHECK_CUDA(insmos_ensure_smem(k_spconv_tc4<NT, KSC>, smem, configured));
    k_spconv_tc4<NT, KSC><<<(unsigned)ceil_div64(units, nwarps / wpt), nwarps * 32, smem, st>>>(a);
    INSMOS_CHECK_LAUNCH("k_spconv_tc4");
    return INSMOS_OK;
}

template <int NT, int KSC>
static int launch_tc34(const TcArgs& a, cudaStream_t st) {
    static const bool v3 = getenv("INSMOS_TC_V3") != nullptr;         // A/B switch
    if (v3) return launch_tc3<NT, KSC>(a, st);
    return launch_tc4<NT, KSC>(a, st);
}

template <int NT>
static int dispatch_ks(const TcArgs& a, cudaStream_t st) {
    if (a.Cin % 8 == 0) {
        switch (a.Cin / 8) {
            case 1: return launch_tc34<NT, 1>(a, st);
            case 2: return launch_tc34<NT, 2>(a, st);
            case 3: return launch_tc34<NT, 3>(a, st);
            case 4: return launch_tc34<NT, 4>(a, st);
            case 6: return launch_tc34<NT, 6>(a, st);
            default: break;
        }
    }
    return launch_tc3<NT, 0>(a, st);
}

extern "C" int insmos_sparse_conv_fwd_tc(const float* in, int64_t n_in, int32_t Cin,
                                         const void* wfrag, int32_t K, int32_t Cout,
                                         const uint16_t* seg, const uint32_t* entries, int32_t TM,
                                         float* out, int64_t n_out,
                                         const insmos_epilogue_t* ep_in, void* stream) {
    if ((n_in > 0 && !in) || !wfrag || !seg || !entries || (n_out > 0 && !out) || Cin <= 0 || Cout <= 0 || K <= 0 || n_out < 0 || n_in < 0)
        return INSMOS_ERR_INVALID_ARG;
    if (TM != 16 && TM != 32 && TM != 64 && TM != 128) return INSMOS_ERR_INVALID_ARG;
    if (n_in > (int64_t)INSMOS_ROW_MASK) return INSMOS_ERR_UNSUPPORTED;       // entry 0xffffffff is the v4 'no pair' marker
    TcArgs a;
    a.in = in; a.wf = (const uint4*)wfrag; a.seg = seg; a.entries = entries; a.out = out;
    a.n_out = n_out; a.n_tiles = ceil_div64(n_out, TM);
    a.Cin = Cin; a.Cout = Cout; a.K = K; a.TM = TM; a.KS = (Cin + 7) / 8; a.NT8 = (Cout + 7) / 8;
    a.ep = insmos_epilogue_t{nullptr, nullptr, nullptr, nullptr, 0};
    if (ep_in) a.ep = *ep_in;
    if (a.ep.scale && !a.ep.shift) return INSMOS_ERR_INVALID_ARG;
    if (n_out == 0) return INSMOS_OK;
    a.groups = a.NT8;
    a.terms = 3;                              // INSMOS_TF32_TERMS=1|2: accuracy study of cheaper splits (tests/accuracy_tf32_terms.py), never the product default
    if (const char* e = getenv("INSMOS_TF32_TERMS")) { const int v = atoi(e); if (v == 1 || v == 2) a.terms = v; }
    // >= 64 output channels: weight traffic dominates -> block-cooperative kernel with the slice's weights in smem
    // (measured on B200, C2 workload: 128->128 K=27 430 -> 283 us, 256->128 843 -> 494 us; at 32 channels the
    // per-bucket barriers cost more than the saved traffic: 48->32 K=81 240 -> 550 us)
    if (a.NT8 >= 8 && getenv("INSMOS_NO_BIG") == nullptr) return launch_tc_big(a, (cudaStream_t)stream);
    // two n-tiles per warp halve the redundant gathers; only when that still leaves thousands of warps
    static const bool nt2_always = getenv("INSMOS_TC_NT1") == nullptr;    // several warps per tile provide the parallelism: two n-tiles per warp whenever possible (A/B: 1.725 -> 1.70 ms)
    if (a.NT8 % 2 == 0 && (nt2_always || a.n_tiles * (a.NT8 / 2) >= 4096)) {
        a.groups = a.NT8 / 2;
        return dispatch_ks<2>(a, (cudaStream_t)stream);
    }
    a.groups = a.NT8;
    return dispatch_ks<1>(a, (cudaStream_t)stream);
}
