// Sparse convolution on the 5th-generation tensor cores (tcgen05.mma kind::tf32, TMEM accumulators, TMA bulk copies of
// the weight tiles, mbarrier pipeline) -- the path for the wide layers of the 3D U-Net (SURVEY.md 8a rows a8/a12:
// SubMConv3d / SparseConv3d / SparseInverseConv3d with 32..256 channels on 6 k .. 42 k rows), where the legacy
// mma.sync kernels are bound by re-fetching the [Cin x Cout] weight slice per 16-pair chunk.
//
// Formulation: OUTPUT-STATIONARY implicit GEMM.  A CTA owns a super-tile of 128 consecutive output rows (G = 128/TM
// rule-book tiles) and NB output channels.  For every kernel offset k that has at least one pair in the super-tile and
// every chunk of 32 input channels, the A operand is the [128 x 32] matrix whose row r is the feature row of output
// row r's neighbour at offset k (zeros when the neighbour is absent), the B operand is W[k][chunk, :NB].  All offsets
// accumulate into the SAME TMEM tile D[128 x NB]: no scatter, no atomics, one epilogue (BN scale/shift, bias, residual,
// ReLU) that writes every output row once.  Unoccupied (row, k) slots cost MMA rows of zeros -- at 3^3 kernels on
// these levels 40-60 % of the slots are occupied, and the tensor pipe has ~8x the mma.sync throughput.
// Same arithmetic contract as conv_tc.cu: 3xTF32 (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi), fp32 accumulate.
//
// Roles (10 warps):
//   warps 0-7  gather producers: 8 lanes per row (one 128-byte line per row and channel chunk, coalesced), TF32 hi/lo
//              split, stores into the K-major SWIZZLE_128B operand images, fence.proxy.async + mbarrier arrive; the
//              loads of chunk c+1 are issued before chunk c is split and stored.  Afterwards: epilogue from TMEM.
//   warp 8     TMA: two cp.async.bulk per stage (hi and lo rows [n0, n0+NB) of the pre-swizzled weight image).
//   warp 9     TMEM allocation; lane 0 issues the MMAs and commits stages / the accumulator.
#include "umma.cuh"
#include <stdlib.h>

#define UM_BM 128
#define UM_BK 32
#define UM_PROD_WARPS 8
#define UM_PROD_THREADS (UM_PROD_WARPS * 32)
#define UM_THREADS (UM_PROD_THREADS + 64)
#define UM_A_BYTES (UM_BM * 128)                    // one [128 x 32] fp32 operand image
#define UM_MAX_STAGES 4

struct UmArgs {
    const float* in; const float* wimg; const uint16_t* seg; const uint32_t* entries; float* out;
    int64_t n_out, n_tiles;
    int Cin, Cout, K, TM, G, cchunks, stages, nsplit, NB, nacc;
    insmos_epilogue_t ep;
};

// acc*scale + shift (+ bias) (+ residual) (relu) for 4 consecutive channels; scale/shift/bias come from the CTA's shared copy
__device__ __forceinline__ float4 um_epilogue4(float4 o, const float* epc, int NB, int c, int64_t row, const insmos_epilogue_t& ep) {
    const float4 sc = *reinterpret_cast<const float4*>(epc + c);
    const float4 sh = *reinterpret_cast<const float4*>(epc + NB + c);
    const float4 bi = *reinterpret_cast<const float4*>(epc + 2 * NB + c);
    if (ep.scale) { o.x = __fmaf_rn(o.x, sc.x, sh.x); o.y = __fmaf_rn(o.y, sc.y, sh.y); o.z = __fmaf_rn(o.z, sc.z, sh.z); o.w = __fmaf_rn(o.w, sc.w, sh.w); }
    if (ep.bias) { o.x += bi.x; o.y += bi.y; o.z += bi.z; o.w += bi.w; }
    if (ep.residual) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(ep.residual + row * NB + c));
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (ep.relu) { o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f); o.w = fmaxf(o.w, 0.0f); }
    return o;
}

// ------------------------------------------------------------------------------------------------
// TS variant: the gathered A operand goes to TENSOR MEMORY instead of shared memory.
// Measured on B200 (tools/micro/umma_rate.cu, profiles/r01_umma_notes.md): a tcgen05.mma kind::tf32 M=128 K=8 occupies
// the tensor pipe for ~75-86 cycles whatever N <= 128 is, i.e. ~1000 cycles per 32-channel chunk with the 3xTF32
// products.  In the first (SS, removed) kernel the MMA operand reads (6-8 KB per instruction), the producers' 32 KB of stores and
// the TMA's weight tiles share the 128 B/clk shared-memory port and a chunk takes ~2000 cycles.  Here the producers
// write the TF32 hi/lo rows straight into TMEM with tcgen05.st (thread = tile row = TMEM lane) and only the weight
// tile is read from shared memory.  tcgen05.wait::st does not wait for global loads, so the next chunk's gathers stay
// in flight in registers while the current one is split and stored.
// TMEM map (512 columns): [0,256) four A stages (hi 32 | lo 32 columns each); [256, 256 + nacc*NB) accumulators.
// Layers with few super-tiles split the active (offset, chunk) list over `ksplit` CTAs; the partial tiles go through
// an L2-resident scratch buffer and the LAST CTA of a super-tile (atomic ticket) adds them in split order -- the
// summation order is fixed, so results do not depend on scheduling -- and applies the epilogue.
#define UT_STAGES 4

struct UtArgs {
    const float* in; const float* wimg; const uint16_t* seg; const uint32_t* entries; float* out;
    float* partial; unsigned int* counters;
    int64_t n_out, n_tiles, n_pad;
    int Cin, Cout, K, TM, G, cchunks, nacc, ksplit, stages, tmem_cols;
    int dense_H, dense_W;                                            // > 0: dense 3x3 pad-1 image convolution (rows = pixels), no rule book
    insmos_epilogue_t ep;
};

__global__ void __launch_bounds__(UM_THREADS)
k_spconv_umma_ts(UtArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int K = p.K, TM = p.TM, G = p.G, NB = p.Cout, Cin = p.Cin, nacc = p.nacc;
    const int S = p.stages, slog = (S == 4) ? 2 : 1;                 // 2 or 4 stages (A in TMEM, B in shared memory)
    const int acc_col = S * 64;                                      // accumulators follow the A stages
    const int b_bytes = NB * 128;                                    // one [NB x 32] weight image
    const int stage_bytes = 2 * b_bytes;                             // B hi | B lo
    int* nbr = reinterpret_cast<int*>(smem + (size_t)S * stage_bytes);   // [K][128]: in_row + 1, 0 = absent
    int* klist = nbr + K * UM_BM;
    int* meta = klist + K;                                           // [0] active offsets, [1] "last CTA" flag
    uint64_t* bars = reinterpret_cast<uint64_t*>(meta + 2 + ((K * (UM_BM + 1)) & 1));   // full[4], empty[4], accumulator ready
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * UT_STAGES + 1);
    uint16_t* sseg = reinterpret_cast<uint16_t*>(tmem_slot + 2);     // [G][K+1] copy of the tiles' bucket offsets
    // epilogue constants (scale | shift | bias, NB floats each) staged once per CTA: the epilogue reads them as broadcast
    // LDS.128 instead of three scalar global loads per output element (ncu: 10 % of the dense kernel's stall samples)
    float* epc = reinterpret_cast<float*>(smem + (((size_t)(reinterpret_cast<uint8_t*>(sseg + (size_t)G * (K + 1)) - smem) + 15) & ~(size_t)15));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t stile = blockIdx.x / p.ksplit;
    const int sp = (int)(blockIdx.x - stile * p.ksplit);
    const int64_t tile0 = stile * G;
    const int ntile = (int)((p.n_tiles - tile0) < G ? (p.n_tiles - tile0) : G);

    if (tid == 0) {
        for (int s = 0; s < UT_STAGES; ++s) {
            mbar_init(smem_u32(bars + s), 4 + 1);                        // 4 producer warps of the stage's group + the TMA expect_tx
            mbar_init(smem_u32(bars + UT_STAGES + s), 1);                // released by tcgen05.commit
        }
        mbar_init(smem_u32(bars + 2 * UT_STAGES), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int c = threadIdx.x; c < NB; c += UM_THREADS) {
        epc[c] = p.ep.scale ? __ldg(p.ep.scale + c) : 1.0f;
        epc[NB + c] = p.ep.scale ? __ldg(p.ep.shift + c) : 0.0f;
        epc[2 * NB + c] = p.ep.bias ? __ldg(p.ep.bias + c) : 0.0f;
    }
    if (warp == UM_PROD_WARPS + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    const bool dense = p.dense_W > 0;
    if (dense) {
        // dense image mode (BEV 3x3 convolutions): the neighbour of pixel m at tap k is pixel m + dy*W + dx when in bounds
        const int HW = p.dense_H * p.dense_W;
        for (int i = tid; i < K * UM_BM; i += UM_THREADS) {
            const int k = i / UM_BM, r = i - k * UM_BM;
            const int64_t m = tile0 * TM + r;
            int v = 0;
            if (m < HW) {
                const int py = (int)(m / p.dense_W), px = (int)(m - (int64_t)py * p.dense_W);
                const int y = py + k / 3 - 1, x = px + k % 3 - 1;
                if ((unsigned)y < (unsigned)p.dense_H && (unsigned)x < (unsigned)p.dense_W) v = y * p.dense_W + x + 1;
            }
            nbr[i] = v;
        }
        for (int k = tid; k < K; k += UM_THREADS) klist[k] = k;
        if (tid == 0) { meta[0] = K; meta[1] = 0; }
    } else {
    for (int i = tid; i < K * UM_BM; i += UM_THREADS) nbr[i] = 0;
    for (int i = tid; i < G * (K + 1); i += UM_THREADS) {
        const int gi = i / (K + 1);
        sseg[i] = (gi < ntile) ? p.seg[(tile0 + gi) * (K + 1) + (i - gi * (K + 1))] : (uint16_t)0;
    }
    __syncthreads();
    // dense neighbour table: one thread per rule-book entry (all loads independent), its bucket by binary search
    for (int gi = 0; gi < ntile; ++gi) {
        const uint16_t* ts = sseg + gi * (K + 1);
        const int tot = ts[K];
        const uint32_t* tent = p.entries + (tile0 + gi) * (int64_t)TM * K;
        for (int e = tid; e < tot; e += UM_THREADS) {
            const uint32_t ent = __ldg(tent + e);
            int lo = 0, hi = K;                                          // largest k with ts[k] <= e
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((int)ts[mid] <= e) lo = mid; else hi = mid; }
            nbr[lo * UM_BM + gi * TM + (int)(ent >> INSMOS_ROW_BITS)] = (int)(ent & INSMOS_ROW_MASK) + 1;
        }
    }
    }
    if (warp == 0 && !dense) {
        int cnt = 0;
        for (int k0 = 0; k0 < K; k0 += 32) {
            const int k = k0 + lane;
            int tot = 0;
            if (k < K)
                for (int gi = 0; gi < ntile; ++gi) tot += (int)sseg[gi * (K + 1) + k + 1] - (int)sseg[gi * (K + 1) + k];
            const unsigned bal = __ballot_sync(0xffffffffu, tot > 0);
            if (tot > 0) klist[cnt + __popc(bal & ((1u << lane) - 1u))] = k;
            cnt += __popc(bal);
        }
        if (lane == 0) { meta[0] = cnt; meta[1] = 0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const int cchunks = p.cchunks;
    const int ntotal = meta[0] * cchunks;
    const int c_begin = (int)(((int64_t)sp * ntotal) / p.ksplit), c_end = (int)(((int64_t)(sp + 1) * ntotal) / p.ksplit);
    const int nloc = c_end - c_begin;                                // chunks of this CTA
    const int kidx0 = c_begin / cchunks, kc0 = c_begin - kidx0 * cchunks;

    if (warp < UM_PROD_WARPS) {
        // ---------------- gather producers: thread = tile row = TMEM lane; group g takes local chunks g, g+2, ... ----------------
        const int grp = warp >> 2;
        const int r = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const bool vec = (Cin & 3) == 0;
        const float* __restrict__ in = p.in;
        float4 v[8];
        auto issue = [&](int kidx, int kc) {
            const int src = nbr[klist[kidx] * UM_BM + r];
            const int c0 = kc * UM_BK;
            const int cw = min(UM_BK, Cin - c0);
            const float* x = in + (size_t)(src > 0 ? src - 1 : 0) * Cin + c0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (src > 0 && 4 * j < cw) {
                    if (vec) v[j] = __ldg(reinterpret_cast<const float4*>(x + 4 * j));
                    else {
                        v[j].x = __ldg(x + 4 * j);
                        if (4 * j + 1 < cw) v[j].y = __ldg(x + 4 * j + 1);
                        if (4 * j + 2 < cw) v[j].z = __ldg(x + 4 * j + 2);
                        if (4 * j + 3 < cw) v[j].w = __ldg(x + 4 * j + 3);
                    }
                }
            }
        };
        int kidx = kidx0, kc = kc0 + grp;                                // (offset index, channel chunk) of local chunk i
        while (kc >= cchunks) { kc -= cchunks; ++kidx; }
        if (grp < nloc) issue(kidx, kc);
        for (int i = grp; i < nloc; i += 2) {
            const int s = i & (S - 1);
            const uint32_t ph = (uint32_t)(i >> slog) & 1u;
            const int cw = min(UM_BK, Cin - kc * UM_BK);
            const int ksteps = (cw + 7) >> 3;
            float4 cur[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) cur[j] = v[j];
            kc += 2;
            while (kc >= cchunks) { kc -= cchunks; ++kidx; }
            if (i + 2 < nloc) issue(kidx, kc);                           // next chunk's gathers stay in flight
            mbar_wait(smem_u32(bars + UT_STAGES + s), ph ^ 1u);          // A columns and B tile of the stage are free
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = lane_addr + (uint32_t)(s * 64), a_lo = a_hi + 32u;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h * 2 < ksteps) {                                    // warp-uniform: tcgen05.st is warp-collective
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 x = cur[h * 4 + j];
                        float a, b;
                        split_rn(x.x, a, b); hi[4 * j + 0] = __float_as_uint(a); lo[4 * j + 0] = __float_as_uint(b);
                        split_rn(x.y, a, b); hi[4 * j + 1] = __float_as_uint(a); lo[4 * j + 1] = __float_as_uint(b);
                        split_rn(x.z, a, b); hi[4 * j + 2] = __float_as_uint(a); lo[4 * j + 2] = __float_as_uint(b);
                        split_rn(x.w, a, b); hi[4 * j + 3] = __float_as_uint(a); lo[4 * j + 3] = __float_as_uint(b);
                    }
                    tmem_st16(a_hi + (uint32_t)(h * 16), hi);
                    tmem_st16(a_lo + (uint32_t)(h * 16), lo);
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(bars + s));
        }
        // ---------------- epilogue ----------------
        if (nloc > 0) {
            mbar_wait(smem_u32(bars + 2 * UT_STAGES), 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const int halves = (NB % 32) == 0 ? 2 : 1;                       // warps 4-7 take the upper half of the columns
        const int hcols = NB / halves;
        const int half = warp >> 2;
        const int64_t row = tile0 * TM + r;
        const uint32_t taddr = lane_addr + (uint32_t)(acc_col + half * hcols);
        const bool mine = half < halves;
        const bool direct = p.ksplit == 1;
        if (mine) {
#pragma unroll 1
            for (int cb = 0; cb < hcols; cb += 16) {
                uint32_t rr[16];
                if (nloc > 0) tmem_ld16(taddr + (uint32_t)cb, rr);
                else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) rr[j] = 0u;
                }
                for (int a = 1; a < nacc && a <= nloc; ++a) {            // accumulators that received products
                    uint32_t r2[16];
                    tmem_ld16(taddr + (uint32_t)(a * NB + cb), r2);
#pragma unroll
                    for (int j = 0; j < 16; ++j) rr[j] = __float_as_uint(__uint_as_float(rr[j]) + __uint_as_float(r2[j]));
                }
                if (row < p.n_out) {
                    const int cbase = half * hcols + cb;
                    float* dst = direct ? p.out + row * NB + cbase : p.partial + ((int64_t)sp * p.n_pad + row) * NB + cbase;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 o = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]),
                                               __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
                        if (direct) o = um_epilogue4(o, epc, NB, cbase + 4 * j, row, p.ep);
                        *reinterpret_cast<float4*>(dst + 4 * j) = o;
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (!direct) {
            // ticket: the CTA that arrives last owns the reduction of the super-tile's ksplit partial tiles
            __threadfence();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid == 0) {
                const unsigned int old = atomicAdd(p.counters + stile, 1u);
                meta[1] = (old == (unsigned int)(p.ksplit - 1)) ? 1 : 0;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (*reinterpret_cast<volatile int*>(meta + 1) && mine && row < p.n_out) {
                __threadfence();
#pragma unroll 1
                for (int cb = 0; cb < hcols; cb += 4) {
                    const int cbase = half * hcols + cb;
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int q = 0; q < p.ksplit; ++q) {                 // fixed order: scheduling-independent result
                        const float4 t = __ldcg(reinterpret_cast<const float4*>(p.partial + ((int64_t)q * p.n_pad + row) * NB + cbase));
                        o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
                    }
                    o = um_epilogue4(o, epc, NB, cbase, row, p.ep);
                    *reinterpret_cast<float4*>(p.out + row * NB + cbase) = o;
                }
            }
        }
    } else if (warp == UM_PROD_WARPS) {
        // ---------------- TMA: weight images ----------------
        if (lane == 0) {
            const size_t img_floats = (size_t)NB * 32;
            int kidx = kidx0, kc = kc0;
            for (int i = 0; i < nloc; ++i) {
                const int s = i & (S - 1);
                const uint32_t ph = (uint32_t)(i >> slog) & 1u;
                mbar_wait(smem_u32(bars + UT_STAGES + s), ph ^ 1u);
                const float* hi = p.wimg + ((size_t)klist[kidx] * cchunks + kc) * 2 * img_floats;
                const uint32_t dst = smem_u32(smem + (size_t)s * stage_bytes);
                mbar_arrive_expect_tx(smem_u32(bars + s), 2u * (uint32_t)b_bytes);
                tma_bulk_g2s(dst, hi, 2u * (uint32_t)b_bytes, smem_u32(bars + s));      // hi and lo images are adjacent
                if (++kc == cchunks) { kc = 0; ++kidx; }
            }
        }
    } else {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32_m128(NB);
            int kc = kc0, acc = 0;
            for (int i = 0; i < nloc; ++i) {
                const int s = i & (S - 1);
                const uint32_t ph = (uint32_t)(i >> slog) & 1u;
                mbar_wait(smem_u32(bars + s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t b_hi = umma_desc_sw128(base), b_lo = umma_desc_sw128(base + b_bytes);
                const uint32_t a_hi = tmem_base + (uint32_t)(s * 64), a_lo = a_hi + 32u;
                const int cw = min(UM_BK, Cin - kc * UM_BK);
                const int ksteps = (cw + 7) >> 3;
                // accumulator 0 takes the two small cross products (|a_lo*b_hi|, |a_hi*b_lo| ~ 2^-11 |a*b|), accumulators
                // 1..nacc-1 take the a_hi*b_hi products round-robin: the tensor core truncates on every accumulate, so the
                // drift is proportional to the number of adds into a LARGE accumulator (tests/accuracy_umma.py)
                const uint32_t dsmall = tmem_base + (uint32_t)acc_col;
                const uint32_t dbig = tmem_base + (uint32_t)(acc_col + (nacc > 1 ? 1 + acc : 0) * NB);
                const bool big_first = (nacc > 1) ? (i < nacc - 1) : false;
                for (int j = 0; j < ksteps; ++j) {
                    const uint64_t adv = (uint64_t)(j * 2);
                    umma_tf32_ts(dsmall, a_lo + (uint32_t)(8 * j), b_hi + adv, idesc, (i != 0 || j != 0) ? 1u : 0u);
                    umma_tf32_ts(dsmall, a_hi + (uint32_t)(8 * j), b_lo + adv, idesc, 1u);
                    umma_tf32_ts(dbig, a_hi + (uint32_t)(8 * j), b_hi + adv, idesc, (big_first && j == 0) ? 0u : 1u);
                }
                umma_commit(smem_u32(bars + UT_STAGES + s));
                if (++kc == cchunks) kc = 0;
                if (++acc >= nacc - 1) acc = 0;
            }
            if (nloc > 0) umma_commit(smem_u32(bars + 2 * UT_STAGES));
        }
    }
    __syncthreads();
    if (warp == UM_PROD_WARPS + 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// weight [K, Cin, Cout] fp32 -> per (k, 32-channel chunk): hi image [Cout x 32] then lo image [Cout x 32], K-major
// SWIZZLE_128B: element (n, c) at byte n*128 + (((c>>2) ^ (n&7))<<4) + (c&3)*4; zeros beyond Cin.  Any row range
// [n0, n0+NB) with n0 % 8 == 0 of an image is itself a valid operand tile (the swizzle is row-local).
__global__ void k_umma_prep_weights(const float* __restrict__ w, int K, int Cin, int Cout, int cchunks, float* __restrict__ img) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)K * cchunks * UM_BK * Cout;
    if (idx >= total) return;
    const int n = (int)(idx % Cout);
    const int c = (int)((idx / Cout) % UM_BK);
    const int kc = (int)((idx / ((int64_t)Cout * UM_BK)) % cchunks);
    const int k = (int)(idx / ((int64_t)Cout * UM_BK * cchunks));
    const int ci = kc * UM_BK + c;
    float hi = 0.f, lo = 0.f;
    if (ci < Cin) {
        split_rn(w[((int64_t)k * Cin + ci) * Cout + n], hi, lo);
        lo = __uint_as_float((__float_as_uint(lo) + 0x1000u) & 0xffffe000u);
    }
    float* blk = img + ((size_t)k * cchunks + kc) * 2 * ((size_t)Cout * 32);
    const int off = (n * 128 + ((((c >> 2) ^ (n & 7))) << 4) + (c & 3) * 4) / 4;
    blk[off] = hi;
    blk[(size_t)Cout * 32 + off] = lo;
}

static bool umma_shape_ok(int32_t K, int32_t Cin, int32_t Cout) {
    return K > 0 && K <= 128 && Cin > 0 && Cin <= 1024 && Cout >= 16 && Cout <= 128 && (Cout % 16) == 0;
}

extern "C" int64_t insmos_conv_wimg_elems(int32_t K, int32_t Cin, int32_t Cout) {
    if (!umma_shape_ok(K, Cin, Cout)) return 0;
    return (int64_t)K * ((Cin + UM_BK - 1) / UM_BK) * 2 * Cout * 32;
}

extern "C" int insmos_conv_prep_weights_umma(const float* weight, int32_t K, int32_t Cin, int32_t Cout, float* wimg, void* stream) {
    if (!weight || !wimg) return INSMOS_ERR_INVALID_ARG;
    if (!umma_shape_ok(K, Cin, Cout)) return INSMOS_ERR_UNSUPPORTED;
    const int cchunks = (Cin + UM_BK - 1) / UM_BK;
    const int64_t total = (int64_t)K * cchunks * UM_BK * Cout;
    k_umma_prep_weights<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(weight, K, Cin, Cout, cchunks, wimg);
    INSMOS_CHECK_LAUNCH("k_umma_prep_weights");
    return INSMOS_OK;
}

// launch configuration of the TS kernel (shared by the sparse and the dense-image entry points)
static int launch_umma_ts(UtArgs& t, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
    const int Cout = t.Cout, K = t.K;
    const struct { int64_t n_tiles; int G; int cchunks; } a = {t.n_tiles, t.G, t.cchunks};
    const int64_t st = ceil_div64(a.n_tiles, a.G);
    t.n_pad = st * UM_BM;
    // TMEM budget: S A-stages of 64 columns + nacc accumulators of Cout columns, power of two.  Cout <= 64 fits in 256
    // columns with 2 stages, so two CTAs share an SM and overlap each other's prologue / epilogue; Cout = 128 takes all 512.
    t.stages = Cout <= 64 ? 2 : 4;
    if (const char* e = getenv("INSMOS_UMMA_STAGES")) { const int v = atoi(e); if (v == 2 || v == 4) t.stages = v; }
    const int acc_budget = (t.stages == 2 ? 256 : 512) - t.stages * 64;
    t.nacc = 4;
    while (t.nacc > 1 && Cout * t.nacc > acc_budget) t.nacc /= 2;
    if (const char* e = getenv("INSMOS_UMMA_NACC")) { const int v = atoi(e); if ((v == 1 || v == 2 || v == 4) && Cout * v <= acc_budget) t.nacc = v; }
    t.tmem_cols = 32;
    while (t.tmem_cols < t.stages * 64 + Cout * t.nacc) t.tmem_cols <<= 1;
    const size_t smem_ts = 1024 + (size_t)t.stages * 2 * Cout * 128 + sizeof(int) * ((size_t)K * UM_BM + K + 4) +
                           8 * (2 * UT_STAGES + 1) + 8 + 16 + 2 * (size_t)a.G * (K + 1) + 32 + 3 * sizeof(float) * (size_t)Cout;
    const int per_sm = (t.tmem_cols <= 256 && 2 * (smem_ts + 1024) <= 227 * 1024) ? 2 : 1;
    // offsets split over CTAs while the layer cannot fill the machine, as long as every CTA keeps >= 8 chunks
    int ksplit = (int)((148 * per_sm) / st);
    if (ksplit > 4) ksplit = 4;
    if (ksplit > (K * a.cchunks) / 8) ksplit = (K * a.cchunks) / 8;
    if (ksplit < 1) ksplit = 1;
    if (const char* e = getenv("INSMOS_UMMA_KSPLIT")) { const int v = atoi(e); if (v >= 1 && v <= 4) ksplit = v; }
    const int64_t need = 256 + st * 4 + (int64_t)ksplit * t.n_pad * Cout * 4;
    if (ksplit > 1 && (!workspace || workspace_bytes < need)) ksplit = 1;
    t.ksplit = ksplit;
    t.counters = nullptr; t.partial = nullptr;
    if (ksplit > 1) {
        t.counters = reinterpret_cast<unsigned int*>(workspace);
        t.partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + ((st * 4 + 255) / 256) * 256);
        INSMOS_CHECK_CUDA(cudaMemsetAsync(workspace, 0, (size_t)st * 4, stream));
    }
    if (smem_ts <= 227 * 1024) {
        static thread_local insmos_smem_cfg_t configured_ts;
        INSMOS_CHECK_CUDA(insmos_ensure_smem(k_spconv_umma_ts, smem_ts, configured_ts));
        k_spconv_umma_ts<<<(unsigned)(st * ksplit), UM_THREADS, smem_ts, stream>>>(t);
        INSMOS_CHECK_LAUNCH("k_spconv_umma_ts");
        return INSMOS_OK;
    }
    return INSMOS_ERR_UNSUPPORTED;
}

extern "C" int64_t insmos_sparse_conv_umma_workspace_bytes(int64_t n_out, int32_t Cout) {
    if (n_out <= 0 || Cout <= 0) return 0;
    const int64_t stiles = ceil_div64(n_out, UM_BM);
    if (stiles >= 148) return 0;                                      // enough super-tiles: no offset split
    return 256 + stiles * 4 + 4 * stiles * UM_BM * (int64_t)Cout * 4;
}

extern "C" int insmos_sparse_conv_fwd_umma(const float* in, int64_t n_in, int32_t Cin,
                                           const float* wimg, int32_t K, int32_t Cout,
                                           const uint16_t* seg, const uint32_t* entries, int32_t TM,
                                           float* out, int64_t n_out,
                                           const insmos_epilogue_t* ep_in,
                                           void* workspace, int64_t workspace_bytes, void* stream) {
    if ((n_in > 0 && !in) || !wimg || !seg || !entries || (n_out > 0 && !out) || n_out < 0 || n_in < 0) return INSMOS_ERR_INVALID_ARG;
    if (TM != 16 && TM != 32 && TM != 64 && TM != 128) return INSMOS_ERR_INVALID_ARG;
    if (!umma_shape_ok(K, Cin, Cout)) return INSMOS_ERR_UNSUPPORTED;
    if (n_in > (int64_t)INSMOS_ROW_MASK) return INSMOS_ERR_UNSUPPORTED;
    UmArgs a;
    a.in = in; a.wimg = wimg; a.seg = seg; a.entries = entries; a.out = out;
    a.n_out = n_out; a.n_tiles = ceil_div64(n_out, TM);
    a.Cin = Cin; a.Cout = Cout; a.K = K; a.TM = TM; a.G = UM_BM / TM; a.cchunks = (Cin + UM_BK - 1) / UM_BK;
    a.ep = insmos_epilogue_t{nullptr, nullptr, nullptr, nullptr, 0};
    if (ep_in) a.ep = *ep_in;
    if (a.ep.scale && !a.ep.shift) return INSMOS_ERR_INVALID_ARG;
    if (n_out == 0) return INSMOS_OK;
    if (Cout > 128) return INSMOS_ERR_UNSUPPORTED;
    UtArgs t;
    t.in = in; t.wimg = wimg; t.seg = seg; t.entries = entries; t.out = out;
    t.n_out = n_out; t.n_tiles = a.n_tiles;
    t.Cin = Cin; t.Cout = Cout; t.K = K; t.TM = TM; t.G = a.G; t.cchunks = a.cchunks; t.ep = a.ep;
    t.dense_H = 0; t.dense_W = 0;
    return launch_umma_ts(t, workspace, workspace_bytes, (cudaStream_t)stream);
}

// Dense 3x3 / pad 1 image convolution (BEV backbone, base_bev_backbone.py:33-82 with BatchNorm folded) through the same
// kernel: rows = pixels of the channels-last [H*W, Cin] image, the neighbour table is arithmetic instead of a rule book.
// wimg: insmos_conv_prep_weights_umma of the [9, Cin, Cout] weight.
extern "C" int insmos_conv2d_nhwc_umma(const float* in, int32_t H, int32_t W, int32_t Cin,
                                       const float* wimg, int32_t Cout, const float* bias, int32_t relu, float* out,
                                       void* workspace, int64_t workspace_bytes, void* stream) {
    if (!in || !wimg || !out || H <= 0 || W <= 0) return INSMOS_ERR_INVALID_ARG;
    if (!umma_shape_ok(9, Cin, Cout) || Cout > 128 || (int64_t)H * W > (int64_t)INSMOS_ROW_MASK) return INSMOS_ERR_UNSUPPORTED;
    UtArgs t;
    t.in = in; t.wimg = wimg; t.seg = nullptr; t.entries = nullptr; t.out = out;
    t.n_out = (int64_t)H * W; t.TM = UM_BM; t.G = 1; t.n_tiles = ceil_div64(t.n_out, UM_BM);
    t.Cin = Cin; t.Cout = Cout; t.K = 9; t.cchunks = (Cin + UM_BK - 1) / UM_BK;
    t.ep = insmos_epilogue_t{nullptr, nullptr, bias, nullptr, relu};
    t.dense_H = H; t.dense_W = W;
    return launch_umma_ts(t, workspace, workspace_bytes, (cudaStream_t)stream);
}
