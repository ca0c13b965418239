// Sparse convolution, tensor-core path for the NARROW layers (3xTF32 on mma.sync m16n8k8).
//
// Work decomposition: a unit = (tile of TM output rows, group of NT n-tiles); wpt warps share a unit (bucket k -> warp
// k mod wpt, private partial accumulator tiles in shared memory, summed in the epilogue); the pairs of one rule-book
// bucket are packed 16 at a time into the M dimension; weights pre-arranged in fragment order and pre-split into TF32
// hi/lo (insmos_conv_prep_weights in conv.cu).
//   k_spconv_tc4<NT,KSC>: per-warp chunk list + counted loop, Cin = 8*KSC compile-time (8, 16, 24, 32, 48), Cout < 64.
// Other shapes return INSMOS_ERR_UNSUPPORTED (host side: tcgen05 kernel for wide layers, general SIMT kernel otherwise).
// History of the measurements that shaped it (and of the two generations it replaced, removed in round 2):
// profiles/r01_conv_v2_sass_notes.md, profiles/r01_umma_notes.md section 4.
//   * the TF32 hi/lo split of the gathered activations is hi = (x + 0x1000) & 0xffffe000, lo = x - hi (3 instructions; the
//     tensor core ignores the low 13 mantissa bits of lo) instead of cvt.rna.tf32 (emulated, ~4 instr + NaN path);
//   * the weight fragments of a bucket are loaded once per bucket and kept in registers across its chunks;
//   * the chunk's rule-book entries are prefetched two chunks ahead and the gathered rows one chunk ahead, so the two
//     dependent global-load latencies overlap the current chunk's mma + shared-memory accumulate.
#include "common.cuh"
#include <stdlib.h>

#define TC_WARPS 4

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
    const int TM = p.TM, K = p.K;
    // dead-row elimination: a tile whose rows all lie below *first_row is not needed by the caller
    const bool active = tile < (int)p.n_tiles &&
                        !(p.ep.first_row && (int64_t)(tile + 1) * TM <= (int64_t)__ldg(p.ep.first_row));
    const int nbk = (K + wpt - 1) / wpt;                             // buckets of this warp: k = sub, sub + wpt, ...
    const int cap = nbk * (1 + TM / 16);                             // chunk-list capacity per warp
    float* acc = sm + (size_t)warp * TM * CW;
    uint32_t* list = reinterpret_cast<uint32_t*>(sm + (size_t)nwarps * TM * CW) + (size_t)warp * cap;
    // epilogue constants: a thread always writes the same output channel (its element stride wpt*32 is a multiple of CW), so
    // scale / shift / bias are fetched once here and their latency hides behind the main loop (ncu of the 8->8 layer: 15 % of
    // the stall samples sat on these loads when they were issued per element in the epilogue)
    const int my_c = grp * NT * 8 + ((sub * 32 + lane) % CW);
    float ep_scale = 1.0f, ep_shift = 0.0f, ep_bias = 0.0f;
    if (active && my_c < p.Cout) {
        if (p.ep.scale) { ep_scale = __ldg(p.ep.scale + my_c); ep_shift = __ldg(p.ep.shift + my_c); }
        if (p.ep.bias) ep_bias = __ldg(p.ep.bias + my_c);
    }
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
        if (c < p.Cout) {                                            // c == my_c
            if (p.ep.scale) v = __fmaf_rn(v, ep_scale, ep_shift);
            if (p.ep.bias) v += ep_bias;
            if (p.ep.residual) v += __ldg(p.ep.residual + (row0 + r) * p.Cout + c);
            if (p.ep.relu) v = fmaxf(v, 0.0f);
            p.out[(row0 + r) * p.Cout + c] = v;
        }
    }
}

static int choose_wpt(const TcArgs& a, int NT) {
    // (sizing the warps per tile for the ACTIVE tiles of a dead-row launch was measured and changed nothing: 1.29 vs 1.34 ms for
    // the family; it also changes the grouping of the partial sums, so the launches keep the full-grid choice and stay bit-identical)
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
    if (a.K > 127 || (int64_t)a.TM * a.K >= 65536 || units >= (1ll << 30)) return INSMOS_ERR_UNSUPPORTED;
    const int wpt = choose_wpt(a, NT);
    a.wpt = wpt;
    const int nwarps = wpt > TC_WARPS ? wpt : TC_WARPS;
    const int nbk = (a.K + wpt - 1) / wpt;
    const size_t smem = sizeof(float) * (size_t)nwarps * a.TM * NT * 8 + sizeof(uint32_t) * (size_t)nwarps * nbk * (1 + a.TM / 16);
    if (smem > 220 * 1024) return INSMOS_ERR_UNSUPPORTED;
    static thread_local insmos_smem_cfg_t configured;
    INSMOS_CHECK_CUDA(insmos_ensure_smem(k_spconv_tc4<NT, KSC>, smem, configured));
    k_spconv_tc4<NT, KSC><<<(unsigned)ceil_div64(units, nwarps / wpt), nwarps * 32, smem, st>>>(a);
    INSMOS_CHECK_LAUNCH("k_spconv_tc4");
    return INSMOS_OK;
}

template <int NT>
static int dispatch_ks(const TcArgs& a, cudaStream_t st) {
    if (a.Cin % 8 == 0) {
        switch (a.Cin / 8) {
            case 1: return launch_tc4<NT, 1>(a, st);
            case 2: return launch_tc4<NT, 2>(a, st);
            case 3: return launch_tc4<NT, 3>(a, st);
            case 4: return launch_tc4<NT, 4>(a, st);
            case 6: return launch_tc4<NT, 6>(a, st);
            default: break;
        }
    }
    return INSMOS_ERR_UNSUPPORTED;                       // other channel counts: the host side uses the general SIMT kernel
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
    if (a.NT8 >= 8) return INSMOS_ERR_UNSUPPORTED;       // >= 64 output channels belong to the tcgen05 kernel (spconv_umma.cu)
    // two n-tiles per warp halve the redundant gathers; only when that still leaves thousands of warps
    static const bool nt2_always = getenv("INSMOS_TC_NT1") == nullptr;    // several warps per tile provide the parallelism: two n-tiles per warp whenever possible (A/B: 1.725 -> 1.70 ms)
    if (a.NT8 % 2 == 0 && (nt2_always || a.n_tiles * (a.NT8 / 2) >= 4096)) {
        a.groups = a.NT8 / 2;
        return dispatch_ks<2>(a, (cudaStream_t)stream);
    }
    a.groups = a.NT8;
    return dispatch_ks<1>(a, (cudaStream_t)stream);
}
