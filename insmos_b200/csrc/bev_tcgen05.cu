// Dense BEV convolutions on the 5th-generation tensor cores: tcgen05.mma (kind::tf32) with TMEM accumulators,
// TMA bulk copies for the weight tiles and an mbarrier pipeline.  Same arithmetic contract as bev.cu (3xTF32 split:
// a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, fp32 accumulate), same NHWC implicit GEMM:
//   M = 128 output pixels per CTA, N = 128 output channels, K = taps x Cin in chunks of 32 channels (= one 128-byte
//   row of a K-major SWIZZLE_128B tile).
// Roles (192 threads):
//   warps 0-3  A producers: thread r owns pixel row r of the tile: loads the 32 channels of the current tap, splits
//              them into TF32 hi/lo and writes the two 128-byte rows into shared memory in the canonical
//              K-major/SWIZZLE_128B layout (16-byte chunk j of row r at r*128 + ((j ^ (r&7))<<4)); then
//              fence.proxy.async + mbarrier arrive.  Afterwards the same warps are the epilogue: tcgen05.ld the
//              accumulator rows from TMEM, bias (folded BatchNorm shift) + ReLU, 16-byte NHWC stores.
//   warp 4     TMA: one cp.async.bulk per stage brings the pre-swizzled [128 n x 32 k] hi and lo weight images
//              (contiguous 32 KB in global memory, written once by insmos_bev_prep_weights_tcgen05).
//   warp 5     allocates TMEM (128 columns), and its lane 0 issues the MMAs: per stage 4 K-steps x 3 products of
//              tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=128, K=8), then tcgen05.commit to release the stage.
// Descriptor encodings follow cute/arch/mma_sm100_desc.hpp / cute/atom/mma_traits_sm100.hpp (vendored CUTLASS headers):
// K-major SWIZZLE_128B: LBO = 1, SBO = 64 (1024 B between 8-row groups), version = 1, layout_type = 2; the K-step
// inside the 128-byte row advances the start address by 32 bytes.
#include "umma.cuh"

#define T5_BM 128
#define T5_BN 128
#define T5_BK 32
#define T5_STAGES 3
#define T5_TILE_BYTES (T5_BM * 128)                 // one 128x32 fp32 operand image
#define T5_STAGE_BYTES (4 * T5_TILE_BYTES)          // A_hi, A_lo, B_hi, B_lo
#define T5_THREADS 192

struct T5Args {
    const float* in;      // [H*W, Cin] NHWC
    const float* wimg;    // pre-swizzled weight images: [tap][kchunk][nblock] x (hi 16 KB | lo 16 KB)
    const float* bias;    // [Cout] or null
    float* out;
    int H, W, Cin, Cout, mode, relu;
};

__global__ void __launch_bounds__(T5_THREADS, 1)
k_conv_nhwc_tcgen05(T5Args p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: stages (1024-aligned), then barriers, then the TMEM base address slot
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T5_STAGES * T5_STAGE_BYTES);
    // bars[0..S) full, bars[S..2S) empty, bars[2S] accumulator ready
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * T5_STAGES + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int HW = p.H * p.W;
    const int m0 = blockIdx.x * T5_BM;
    const int nblk = blockIdx.y;
    const int ztap = (p.mode == 2) ? blockIdx.z : 0;
    const int ntaps = (p.mode == 0) ? 9 : 1;
    const int cchunks = p.Cin / T5_BK;
    const int nchunks = ntaps * cchunks;
    const int nblocks = p.Cout / T5_BN;

    if (tid == 0) {
        for (int s = 0; s < T5_STAGES; ++s) {
            mbar_init(smem_u32(bars + s), 128 + 1);                      // 128 A-producer threads + the TMA arrive
            mbar_init(smem_u32(bars + T5_STAGES + s), 1);                // released by tcgen05.commit
        }
        mbar_init(smem_u32(bars + 2 * T5_STAGES), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {                                                     // TMEM: 128 fp32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp < 4) {
        // ---------------- A producers ----------------
        const int r = tid;                                               // tile row = pixel m0 + r
        const int m = m0 + r;
        const int py = (m < HW) ? m / p.W : 0, px = (m < HW) ? m % p.W : 0;
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % T5_STAGES;
            const uint32_t ph = (c / T5_STAGES) & 1;
            const int tap = (p.mode == 0) ? c / cchunks : ztap;
            const int c0 = (c % cchunks) * T5_BK;
            const int dy = (p.mode == 0) ? tap / 3 - 1 : 0, dx = (p.mode == 0) ? tap % 3 - 1 : 0;
            float4 v[8];
            const int y = py + dy, x = px + dx;
            const bool inb = (m < HW) && (unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W;
            if (inb) {
                const float4* src = reinterpret_cast<const float4*>(p.in + ((size_t)y * p.W + x) * p.Cin + c0);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __ldg(src + j);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            mbar_wait(smem_u32(bars + T5_STAGES + s), ph ^ 1);           // stage free?
            uint8_t* a_hi = smem + s * T5_STAGE_BYTES;
            uint8_t* a_lo = a_hi + T5_TILE_BYTES;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 h, l;
                split_rn(v[j].x, h.x, l.x); split_rn(v[j].y, h.y, l.y); split_rn(v[j].z, h.z, l.z); split_rn(v[j].w, h.w, l.w);
                const int off = r * 128 + ((j ^ (r & 7)) << 4);
                *reinterpret_cast<float4*>(a_hi + off) = h;
                *reinterpret_cast<float4*>(a_lo + off) = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
            mbar_arrive(smem_u32(bars + s));
        }
        // ---------------- epilogue ----------------
        mbar_wait(smem_u32(bars + 2 * T5_STAGES), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        int64_t orow = m;
        if (p.mode == 2) orow = (int64_t)(2 * py + (ztap >> 1)) * (2 * p.W) + 2 * px + (ztap & 1);
        float* dst = p.out + orow * p.Cout + (size_t)nblk * T5_BN;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int cb = 0; cb < T5_BN; cb += 32) {
            uint32_t rr[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]),
                  "=r"(rr[8]), "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15]),
                  "=r"(rr[16]), "=r"(rr[17]), "=r"(rr[18]), "=r"(rr[19]), "=r"(rr[20]), "=r"(rr[21]), "=r"(rr[22]), "=r"(rr[23]),
                  "=r"(rr[24]), "=r"(rr[25]), "=r"(rr[26]), "=r"(rr[27]), "=r"(rr[28]), "=r"(rr[29]), "=r"(rr[30]), "=r"(rr[31])
                : "r"(taddr + (uint32_t)cb));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (m < HW) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float4 o;
                    o.x = __uint_as_float(rr[4 * q + 0]); o.y = __uint_as_float(rr[4 * q + 1]);
                    o.z = __uint_as_float(rr[4 * q + 2]); o.w = __uint_as_float(rr[4 * q + 3]);
                    if (p.bias) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + nblk * T5_BN + cb + 4 * q));
                        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                    }
                    if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    *reinterpret_cast<float4*>(dst + cb + 4 * q) = o;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else if (warp == 4) {
        // ---------------- TMA: weight images ----------------
        if (lane == 0) {
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % T5_STAGES;
                const uint32_t ph = (c / T5_STAGES) & 1;
                const int tap = (p.mode == 0) ? c / cchunks : ztap;
                const int kc = c % cchunks;
                mbar_wait(smem_u32(bars + T5_STAGES + s), ph ^ 1);
                const float* src = p.wimg + (((size_t)tap * cchunks + kc) * nblocks + nblk) * (2 * T5_TILE_BYTES / 4);
                mbar_arrive_expect_tx(smem_u32(bars + s), 2 * T5_TILE_BYTES);
                tma_bulk_g2s(smem_u32(smem + s * T5_STAGE_BYTES + 2 * T5_TILE_BYTES), src, 2 * T5_TILE_BYTES, smem_u32(bars + s));
            }
        }
    } else {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            // kind::tf32, D = F32, A/B K-major, N = 128, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(T5_BN >> 3) << 17) | ((uint32_t)(T5_BM >> 4) << 24);
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % T5_STAGES;
                const uint32_t ph = (c / T5_STAGES) & 1;
                mbar_wait(smem_u32(bars + s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t base = smem_u32(smem + s * T5_STAGE_BYTES);
                const uint64_t a_hi = umma_desc_sw128(base), a_lo = umma_desc_sw128(base + T5_TILE_BYTES);
                const uint64_t b_hi = umma_desc_sw128(base + 2 * T5_TILE_BYTES), b_lo = umma_desc_sw128(base + 3 * T5_TILE_BYTES);
#pragma unroll
                for (int j = 0; j < T5_BK / 8; ++j) {
                    const uint64_t adv = (uint64_t)(j * 2);                 // 32 bytes per K-step, in 16-byte units
                    umma_tf32(tmem_base, a_lo + adv, b_hi + adv, idesc, (c | j) != 0);
                    umma_tf32(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
                    umma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, 1u);
                }
                umma_commit(smem_u32(bars + T5_STAGES + s));                // stage reusable once these MMAs retire
            }
            umma_commit(smem_u32(bars + 2 * T5_STAGES));                    // accumulator complete
        }
    }
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
    }
}

// weight [taps, Cin, Cout] fp32 (BN folded) -> pre-swizzled TF32 hi/lo images, one 32 KB block per (tap, kchunk, nblock):
// element (n, k) of the [128 x 32] K-major tile at byte n*128 + (((k>>2) ^ (n&7))<<4) + (k&3)*4; hi image then lo image.
__global__ void k_bev_prep_weights(const float* __restrict__ w, int taps, int Cin, int Cout, float* __restrict__ img) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)taps * Cin * Cout;
    if (idx >= total) return;
    const int co = (int)(idx % Cout);
    const int ci = (int)((idx / Cout) % Cin);
    const int tap = (int)(idx / ((int64_t)Cout * Cin));
    const int cchunks = Cin / T5_BK, nblocks = Cout / T5_BN;
    const int kc = ci / T5_BK, k = ci % T5_BK, nb = co / T5_BN, n = co % T5_BN;
    float hi, lo;
    split_rn(w[idx], hi, lo);
    lo = __uint_as_float((__float_as_uint(lo) + 0x1000u) & 0xffffe000u);
    float* blk = img + (((size_t)tap * cchunks + kc) * nblocks + nb) * (2 * T5_TILE_BYTES / 4);
    const int off = (n * 128 + ((((k >> 2) ^ (n & 7))) << 4) + (k & 3) * 4) / 4;
    blk[off] = hi;
    blk[T5_TILE_BYTES / 4 + off] = lo;
}

extern "C" int64_t insmos_bev_wimg_elems(int32_t taps, int32_t Cin, int32_t Cout) {
    return (int64_t)taps * (Cin / T5_BK) * (Cout / T5_BN) * (2 * T5_TILE_BYTES / 4);
}
extern "C" int insmos_bev_prep_weights_tcgen05(const float* weight, int32_t taps, int32_t Cin, int32_t Cout, float* wimg, void* stream) {
    if (!weight || !wimg || taps <= 0 || Cin % T5_BK != 0 || Cout % T5_BN != 0) return INSMOS_ERR_INVALID_ARG;
    const int64_t total = (int64_t)taps * Cin * Cout;
    k_bev_prep_weights<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(weight, taps, Cin, Cout, wimg);
    INSMOS_CHECK_LAUNCH("k_bev_prep_weights");
    return INSMOS_OK;
}

extern "C" int insmos_conv2d_nhwc_tcgen05(const float* in, int32_t H, int32_t W, int32_t Cin,
                                          const float* wimg, int32_t mode, int32_t Cout,
                                          const float* bias, int32_t relu, float* out, void* stream) {
    if (!in || !wimg || !out || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || mode < 0 || mode > 2) return INSMOS_ERR_INVALID_ARG;
    if (Cin % T5_BK != 0 || Cout % T5_BN != 0) return INSMOS_ERR_UNSUPPORTED;
    T5Args p{in, wimg, bias, out, H, W, Cin, Cout, mode, relu};
    constexpr size_t smem = 1024 + (size_t)T5_STAGES * T5_STAGE_BYTES + 8 * (2 * T5_STAGES + 1) + 16;
    static thread_local insmos_smem_cfg_t configured;
    INSMOS_CHECK_CUDA(insmos_ensure_smem(k_conv_nhwc_tcgen05, smem, configured));
    dim3 grid((unsigned)((H * W + T5_BM - 1) / T5_BM), (unsigned)(Cout / T5_BN), mode == 2 ? 4u : 1u);
    k_conv_nhwc_tcgen05<<<grid, T5_THREADS, smem, (cudaStream_t)stream>>>(p);
    INSMOS_CHECK_LAUNCH("k_conv_nhwc_tcgen05");
    return INSMOS_OK;
}
