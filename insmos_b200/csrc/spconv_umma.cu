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

__device__ __forceinline__ float um_epilogue(float v, int c, int64_t row, int Cout, const insmos_epilogue_t& ep) {
    if (ep.scale) v = __fmaf_rn(v, __ldg(ep.scale + c), __ldg(ep.shift + c));
    if (ep.bias) v += __ldg(ep.bias + c);
    if (ep.residual) v += __ldg(ep.residual + row * Cout + c);
    if (ep.relu) v = fmaxf(v, 0.0f);
    return v;
}

__global__ void __launch_bounds__(UM_THREADS)
k_spconv_umma(UmArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int K = p.K, TM = p.TM, G = p.G, NB = p.NB, S = p.stages, Cin = p.Cin;
    const int b_bytes = NB * 128;                                    // one [NB x 32] weight image
    const int stage_bytes = 2 * UM_A_BYTES + 2 * b_bytes;
    int* nbr = reinterpret_cast<int*>(smem + (size_t)S * stage_bytes);   // [K][128]: in_row + 1, 0 = absent
    int* klist = nbr + K * UM_BM;                                    // [K] offsets with pairs in this super-tile
    int* meta = klist + K;                                           // [0] = number of active offsets
    uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(meta + 2) + 7) & ~(uintptr_t)7);   // full[], empty[], accumulator ready
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * UM_MAX_STAGES + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t stile = blockIdx.x / p.nsplit;
    const int split = (int)(blockIdx.x - stile * p.nsplit);
    const int n0 = split * NB;
    const int64_t tile0 = stile * G;
    const int ntile = (int)((p.n_tiles - tile0) < G ? (p.n_tiles - tile0) : G);
    // nacc TMEM accumulators used round-robin over the chunks and summed (fp32, round-to-nearest) in the epilogue: the
    // tensor core truncates when it adds a product into the accumulator, so a single accumulator drifts by ~0.5 ulp
    // per MMA (measured 5e-5 on O(1) outputs after 27 offsets x 64 channels); nacc chains are nacc x shorter.
    const int nacc = p.nacc;
    int tmem_cols = 32;                                              // power of two >= 32
    while (tmem_cols < NB * nacc) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(bars + s), UM_PROD_WARPS + 1);            // one arrive per producer warp + the TMA expect_tx arrive
            mbar_init(smem_u32(bars + UM_MAX_STAGES + s), 1);            // released by tcgen05.commit
        }
        mbar_init(smem_u32(bars + 2 * UM_MAX_STAGES), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == UM_PROD_WARPS + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- dense neighbour table of the super-tile from the bucketed rule book
    for (int i = tid; i < K * UM_BM; i += UM_THREADS) nbr[i] = 0;
    __syncthreads();
    for (int b = tid; b < G * K; b += UM_THREADS) {
        const int gi = b / K, k = b - gi * K;
        if (gi < ntile) {
            const uint16_t* tseg = p.seg + (tile0 + gi) * (K + 1);
            const int s0 = tseg[k], s1 = tseg[k + 1];
            const uint32_t* tent = p.entries + (tile0 + gi) * (int64_t)TM * K;
            for (int e = s0; e < s1; ++e) {
                const uint32_t ent = __ldg(tent + e);
                nbr[k * UM_BM + gi * TM + (int)(ent >> INSMOS_ROW_BITS)] = (int)(ent & INSMOS_ROW_MASK) + 1;
            }
        }
    }
    if (warp == 0) {                                                 // ordered list of non-empty offsets
        int cnt = 0;
        for (int k0 = 0; k0 < K; k0 += 32) {
            const int k = k0 + lane;
            int tot = 0;
            if (k < K)
                for (int gi = 0; gi < ntile; ++gi) {
                    const uint16_t* tseg = p.seg + (tile0 + gi) * (K + 1);
                    tot += (int)tseg[k + 1] - (int)tseg[k];
                }
            const unsigned bal = __ballot_sync(0xffffffffu, tot > 0);
            if (tot > 0) klist[cnt + __popc(bal & ((1u << lane) - 1u))] = k;
            cnt += __popc(bal);
        }
        if (lane == 0) meta[0] = cnt;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const int nact = meta[0];
    const int cchunks = p.cchunks;
    const int nchunks = nact * cchunks;

    if (warp < UM_PROD_WARPS) {
        // ---------------- gather producers ----------------
        const int q = tid & 7;                                           // 16-byte chunk of the 128-byte row image
        const int rb = tid >> 3;                                         // rows rb, rb+32, rb+64, rb+96
        const bool vec = (Cin & 3) == 0;
        const float* __restrict__ in = p.in;
        float4 v[4];
        auto load = [&](int c) {
            const int k = klist[c / cchunks];
            const int c0 = (c % cchunks) * UM_BK + 4 * q;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int src = nbr[k * UM_BM + rb + 32 * i];
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (src > 0 && c0 < Cin) {
                    const float* x = in + (size_t)(src - 1) * Cin + c0;
                    if (vec) v[i] = __ldg(reinterpret_cast<const float4*>(x));
                    else {
                        v[i].x = __ldg(x);
                        if (c0 + 1 < Cin) v[i].y = __ldg(x + 1);
                        if (c0 + 2 < Cin) v[i].z = __ldg(x + 2);
                        if (c0 + 3 < Cin) v[i].w = __ldg(x + 3);
                    }
                }
            }
        };
        if (nchunks > 0) load(0);
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % S;
            const uint32_t ph = (uint32_t)(c / S) & 1u;
            float4 cur[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) cur[i] = v[i];
            if (c + 1 < nchunks) load(c + 1);                            // next chunk's gathers in flight
            mbar_wait(smem_u32(bars + UM_MAX_STAGES + s), ph ^ 1u);      // stage free?
            uint8_t* a_hi = smem + (size_t)s * stage_bytes;
            uint8_t* a_lo = a_hi + UM_A_BYTES;
            const int cw = min(UM_BK, Cin - (c % cchunks) * UM_BK);      // channels of this chunk
            if (q < 2 * ((cw + 7) >> 3)) {                               // 16-byte chunks the issued K-steps read
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = rb + 32 * i;
                    float4 h, l;
                    split_rn(cur[i].x, h.x, l.x); split_rn(cur[i].y, h.y, l.y);
                    split_rn(cur[i].z, h.z, l.z); split_rn(cur[i].w, h.w, l.w);
                    const int off = r * 128 + ((q ^ (r & 7)) << 4);
                    *reinterpret_cast<float4*>(a_hi + off) = h;
                    *reinterpret_cast<float4*>(a_lo + off) = l;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(bars + s));              // one arrival per warp (256 serialized arrivals cost more than the stage)
        }
        // ---------------- epilogue: TMEM -> registers -> BN / residual / ReLU -> global ----------------
        if (nchunks > 0) {
            mbar_wait(smem_u32(bars + 2 * UM_MAX_STAGES), 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const int halves = (NB % 32) == 0 ? 2 : 1;                       // warps 4-7 take the upper half of the columns
        const int hcols = NB / halves;
        const int half = warp >> 2;
        if (half < halves) {
            const int r = (warp & 3) * 32 + lane;                        // TMEM lane = tile row
            const int64_t row = tile0 * TM + r;
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(half * hcols);
#pragma unroll 1
            for (int cb = 0; cb < hcols; cb += 16) {
                uint32_t rr[16];
                if (nchunks > 0) tmem_ld16(taddr + (uint32_t)cb, rr);
                else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) rr[j] = 0u;
                }
                for (int a = 1; a < nacc && a < nchunks; ++a) {          // accumulators that received at least one chunk
                    uint32_t r2[16];
                    tmem_ld16(taddr + (uint32_t)(a * NB + cb), r2);
#pragma unroll
                    for (int j = 0; j < 16; ++j) rr[j] = __float_as_uint(__uint_as_float(rr[j]) + __uint_as_float(r2[j]));
                }
                if (row < p.n_out) {
                    const int cbase = n0 + half * hcols + cb;
                    float* dst = p.out + row * p.Cout + cbase;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 o;
                        o.x = um_epilogue(__uint_as_float(rr[4 * j + 0]), cbase + 4 * j + 0, row, p.Cout, p.ep);
                        o.y = um_epilogue(__uint_as_float(rr[4 * j + 1]), cbase + 4 * j + 1, row, p.Cout, p.ep);
                        o.z = um_epilogue(__uint_as_float(rr[4 * j + 2]), cbase + 4 * j + 2, row, p.Cout, p.ep);
                        o.w = um_epilogue(__uint_as_float(rr[4 * j + 3]), cbase + 4 * j + 3, row, p.Cout, p.ep);
                        *reinterpret_cast<float4*>(dst + 4 * j) = o;
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else if (warp == UM_PROD_WARPS) {
        // ---------------- TMA: weight images ----------------
        if (lane == 0) {
            const size_t img_floats = (size_t)p.Cout * 32;               // one [Cout x 32] image (hi or lo)
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % S;
                const uint32_t ph = (uint32_t)(c / S) & 1u;
                const int k = klist[c / cchunks], kc = c % cchunks;
                mbar_wait(smem_u32(bars + UM_MAX_STAGES + s), ph ^ 1u);
                const float* hi = p.wimg + ((size_t)k * cchunks + kc) * 2 * img_floats + (size_t)n0 * 32;
                const uint32_t dst = smem_u32(smem + (size_t)s * stage_bytes + 2 * UM_A_BYTES);
                mbar_arrive_expect_tx(smem_u32(bars + s), 2u * (uint32_t)b_bytes);
                tma_bulk_g2s(dst, hi, (uint32_t)b_bytes, smem_u32(bars + s));
                tma_bulk_g2s(dst + (uint32_t)b_bytes, hi + img_floats, (uint32_t)b_bytes, smem_u32(bars + s));
            }
        }
    } else {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32_m128(NB);
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % S;
                const uint32_t ph = (uint32_t)(c / S) & 1u;
                mbar_wait(smem_u32(bars + s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t a_hi = umma_desc_sw128(base), a_lo = umma_desc_sw128(base + UM_A_BYTES);
                const uint64_t b_hi = umma_desc_sw128(base + 2 * UM_A_BYTES), b_lo = umma_desc_sw128(base + 2 * UM_A_BYTES + b_bytes);
                const int cw = min(UM_BK, Cin - (c % cchunks) * UM_BK);
                const int ksteps = (cw + 7) >> 3;
                const uint32_t dacc = tmem_base + (uint32_t)((c % nacc) * NB);
                for (int j = 0; j < ksteps; ++j) {
                    const uint64_t adv = (uint64_t)(j * 2);                 // 32 bytes per K-step, in 16-byte units
                    umma_tf32(dacc, a_lo + adv, b_hi + adv, idesc, (c >= nacc || j != 0) ? 1u : 0u);
                    umma_tf32(dacc, a_hi + adv, b_lo + adv, idesc, 1u);
                    umma_tf32(dacc, a_hi + adv, b_hi + adv, idesc, 1u);
                }
                umma_commit(smem_u32(bars + UM_MAX_STAGES + s));            // stage reusable once these MMAs retire
            }
            if (nchunks > 0) umma_commit(smem_u32(bars + 2 * UM_MAX_STAGES));   // accumulator complete
        }
    }
    __syncthreads();
    if (warp == UM_PROD_WARPS + 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
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
    return K > 0 && K <= 128 && Cin > 0 && Cin <= 1024 && Cout >= 16 && Cout <= 256 && (Cout % 16) == 0;
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

extern "C" int insmos_sparse_conv_fwd_umma(const float* in, int64_t n_in, int32_t Cin,
                                           const float* wimg, int32_t K, int32_t Cout,
                                           const uint16_t* seg, const uint32_t* entries, int32_t TM,
                                           float* out, int64_t n_out,
                                           const insmos_epilogue_t* ep_in, void* stream) {
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
    const int64_t stiles = ceil_div64(a.n_tiles, a.G);
    // split the output channels over CTAs while the layer has too few super-tiles to fill 148 SMs (the gather is
    // repeated per split, the MMA work is not)
    int nsplit = 1;
    while (stiles * nsplit < 120 && Cout / (nsplit * 2) >= 32 && (Cout / (nsplit * 2)) % 16 == 0) nsplit *= 2;
    if (const char* e = getenv("INSMOS_UMMA_NSPLIT")) {
        const int v = atoi(e);
        if (v >= 1 && Cout % v == 0 && (Cout / v) % 16 == 0) nsplit = v;
    }
    a.nsplit = nsplit; a.NB = Cout / nsplit;
    if (a.NB > 128) { a.nsplit = Cout / 128; a.NB = 128; if (Cout % 128) return INSMOS_ERR_UNSUPPORTED; }
    const size_t stage_bytes = 2 * (size_t)UM_A_BYTES + 2 * (size_t)a.NB * 128;
    const size_t fixed = 1024 + sizeof(int) * ((size_t)K * UM_BM + K + 2) + 8 * (2 * UM_MAX_STAGES + 1) + 8 + 16;
    int stages = UM_MAX_STAGES;
    while (stages > 2 && fixed + stages * stage_bytes > 220 * 1024) --stages;
    // many CTAs: two resident CTAs per SM (each with a 2-stage ring) overlap each other's gather latency
    if (stiles * a.nsplit > 148 && fixed + 2 * stage_bytes <= 112 * 1024) stages = 2;
    if (const char* e = getenv("INSMOS_UMMA_STAGES")) { const int v = atoi(e); if (v >= 2 && v <= UM_MAX_STAGES) stages = v; }
    const size_t smem = fixed + stages * stage_bytes;
    if (smem > 227 * 1024) return INSMOS_ERR_UNSUPPORTED;
    a.stages = stages;
    const int ctas_per_sm = (int)((227 * 1024) / (smem + 1024)) < 1 ? 1 : (int)((227 * 1024) / (smem + 1024));
    int nacc = 4;                                                     // TMEM: 512 columns per SM shared by the resident CTAs
    while (nacc > 1 && a.NB * nacc * ctas_per_sm > 512) nacc /= 2;
    if (const char* e = getenv("INSMOS_UMMA_NACC")) { const int v = atoi(e); if ((v == 1 || v == 2 || v == 4) && a.NB * v <= 512) nacc = v; }
    a.nacc = nacc;
    static thread_local size_t configured = 0;
    if (smem > configured) {
        INSMOS_CHECK_CUDA(cudaFuncSetAttribute(k_spconv_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    k_spconv_umma<<<(unsigned)(stiles * a.nsplit), UM_THREADS, smem, (cudaStream_t)stream>>>(a);
    INSMOS_CHECK_LAUNCH("k_spconv_umma");
    return INSMOS_OK;
}
