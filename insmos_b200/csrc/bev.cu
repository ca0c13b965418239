// Dense BEV head convolutions: NHWC implicit GEMM on tensor cores.
//
// The reference runs these through cuDNN (base_bev_backbone.py:84-115: 3x3 convs 256->128, 5x 128->128, a 2x2
// stride-2 transposed conv 128->256).  In fp32 cuDNN picks SIMT sgemm kernels (TF32 must stay off for parity),
// ~3 ms per forward on B200.  Here the same arithmetic runs as ONE implicit-GEMM kernel family on tensor cores
// with the 3xTF32 split (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo, fp32 accumulate ~ 2^-22 relative, i.e. fp32-accurate):
//   M = H*W output pixels (block tile 128), N = Cout (block tile 128), K = taps * Cin (chunks of 32 channels).
//   * activations are NHWC [H*W, C] so a K-chunk of one tap is a contiguous 128-byte row per pixel;
//   * both operand tiles go global -> registers -> (split once into TF32 hi/lo) -> shared memory, double buffered,
//     so the mma loop itself contains no conversions; padded row strides make every fragment load conflict-free;
//   * 8 warps as 2(M) x 4(N), warp tile 64x32 = 4x4 mma.m16n8k8 tiles, 64 fp32 accumulators per lane;
//   * BatchNorm(eval) is folded into the weights/bias on the host side once; bias + ReLU in the epilogue.
// mode 0: 3x3, stride 1, zero padding 1.   mode 1: 1x1.   mode 2: 2x2 stride-2 transposed conv (4 independent
// 1x1 GEMMs, grid.z = tap, each scattering to its own output phase of the [2H, 2W] map).
#include "common.cuh"

#define BM 128
#define BN 128
#define BK 32
#define A_STRIDE (BK + 4)      // floats; (4g + t) % 32 distinct for the A fragment pattern
#define B_STRIDE (BN + 8)      // floats; (8t + g) % 32 distinct for the B fragment pattern
#define CONV_THREADS 256

__device__ __forceinline__ void split2(float x, float& hi, float& lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    const float r = x - hi;
    uint32_t l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
    lo = __uint_as_float(l);
}
__device__ __forceinline__ void mma8(float (&d)[4], const float (&a)[4], float b0, float b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                   "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

struct ConvP {
    const float* in;      // [H*W, Cin]
    const float* w;       // [taps, Cin, Cout]
    const float* bias;    // [Cout] or null
    float* out;           // mode 0/1: [H*W, Cout]; mode 2: [2H*2W, Cout]
    int H, W, Cin, Cout, mode, relu;
};

__global__ void __launch_bounds__(CONV_THREADS, 1)
k_conv_nhwc_tc(ConvP p) {
    extern __shared__ __align__(16) float smem[];
    // per stage: A_hi, A_lo [BM][A_STRIDE]; B_hi, B_lo [BK][B_STRIDE]
    constexpr int A_TILE = BM * A_STRIDE, B_TILE = BK * B_STRIDE, STAGE = 2 * A_TILE + 2 * B_TILE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;                  // 2 x 4 warps
    const int HW = p.H * p.W;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int ztap = (p.mode == 2) ? blockIdx.z : 0;
    const int ntaps = (p.mode == 0) ? 9 : 1;
    const int cchunks = p.Cin / BK;
    const int nchunks = ntaps * cchunks;

    // global -> register staging: A: 128 rows x 8 float4; B: 32 rows x 32 float4; 4 float4 each per thread
    float4 ra[4], rb[4];
    auto load_regs = [&](int chunk) {
        const int tap = (p.mode == 0) ? chunk / cchunks : ztap;
        const int c0 = (chunk % cchunks) * BK;
        const int dy = (p.mode == 0) ? tap / 3 - 1 : 0, dx = (p.mode == 0) ? tap % 3 - 1 : 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * CONV_THREADS;          // 0..1023
            const int r = idx >> 3, c4 = idx & 7;
            const int m = m0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < HW) {
                const int y = m / p.W + dy, x = m % p.W + dx;
                if ((unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W)
                    v = __ldg(reinterpret_cast<const float4*>(p.in + ((int64_t)y * p.W + x) * p.Cin + c0) + c4);
            }
            ra[i] = v;
        }
        const float* wb = p.w + ((int64_t)tap * p.Cin + c0) * p.Cout + n0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * CONV_THREADS;
            const int r = idx >> 5, c4 = idx & 31;
            rb[i] = __ldg(reinterpret_cast<const float4*>(wb + (int64_t)r * p.Cout) + c4);
        }
    };
    auto store_smem = [&](int stage) {
        float* Ah = smem + stage * STAGE; float* Al = Ah + A_TILE; float* Bh = Al + A_TILE; float* Bl = Bh + B_TILE;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * CONV_THREADS;
            const int r = idx >> 3, c4 = idx & 7;
            float4 h, l;
            split2(ra[i].x, h.x, l.x); split2(ra[i].y, h.y, l.y); split2(ra[i].z, h.z, l.z); split2(ra[i].w, h.w, l.w);
            *reinterpret_cast<float4*>(Ah + r * A_STRIDE + c4 * 4) = h;
            *reinterpret_cast<float4*>(Al + r * A_STRIDE + c4 * 4) = l;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * CONV_THREADS;
            const int r = idx >> 5, c4 = idx & 31;
            float4 h, l;
            split2(rb[i].x, h.x, l.x); split2(rb[i].y, h.y, l.y); split2(rb[i].z, h.z, l.z); split2(rb[i].w, h.w, l.w);
            *reinterpret_cast<float4*>(Bh + r * B_STRIDE + c4 * 4) = h;
            *reinterpret_cast<float4*>(Bl + r * B_STRIDE + c4 * 4) = l;
        }
    };

    float acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f; }

    load_regs(0);
    store_smem(0);
    __syncthreads();
    for (int c = 0; c < nchunks; ++c) {
        const int stage = c & 1;
        if (c + 1 < nchunks) load_regs(c + 1);               // global loads in flight during the mma block
        const float* Ah = smem + stage * STAGE; const float* Al = Ah + A_TILE;
        const float* Bh = Al + A_TILE; const float* Bl = Bh + B_TILE;
#pragma unroll
        for (int k8 = 0; k8 < BK / 8; ++k8) {
            float bh[4][2], bl[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = wn * 32 + j * 8 + g;
                bh[j][0] = Bh[(k8 * 8 + t) * B_STRIDE + n]; bh[j][1] = Bh[(k8 * 8 + t + 4) * B_STRIDE + n];
                bl[j][0] = Bl[(k8 * 8 + t) * B_STRIDE + n]; bl[j][1] = Bl[(k8 * 8 + t + 4) * B_STRIDE + n];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = wm * 64 + i * 16 + g;
                float ah[4], al[4];
                ah[0] = Ah[r * A_STRIDE + k8 * 8 + t];           ah[1] = Ah[(r + 8) * A_STRIDE + k8 * 8 + t];
                ah[2] = Ah[r * A_STRIDE + k8 * 8 + t + 4];       ah[3] = Ah[(r + 8) * A_STRIDE + k8 * 8 + t + 4];
                al[0] = Al[r * A_STRIDE + k8 * 8 + t];           al[1] = Al[(r + 8) * A_STRIDE + k8 * 8 + t];
                al[2] = Al[r * A_STRIDE + k8 * 8 + t + 4];       al[3] = Al[(r + 8) * A_STRIDE + k8 * 8 + t + 4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    mma8(acc[i][j], al, bh[j][0], bh[j][1]);
                    mma8(acc[i][j], ah, bl[j][0], bl[j][1]);
                    mma8(acc[i][j], ah, bh[j][0], bh[j][1]);
                }
            }
        }
        if (c + 1 < nchunks) store_smem(stage ^ 1);
        __syncthreads();
    }

    // epilogue: bias (folded BatchNorm shift), ReLU, NHWC store (float2 per lane per row)
    const int outW = (p.mode == 2) ? 2 * p.W : p.W;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int m = m0 + wm * 64 + i * 16 + g + h * 8;
            if (m >= HW) continue;
            int64_t orow = m;
            if (p.mode == 2) {
                const int y = m / p.W, x = m % p.W;
                orow = (int64_t)(2 * y + (ztap >> 1)) * outW + 2 * x + (ztap & 1);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + wn * 32 + j * 8 + 2 * t;
                float v0 = acc[i][j][h * 2 + 0], v1 = acc[i][j][h * 2 + 1];
                if (p.bias) { v0 += __ldg(p.bias + n); v1 += __ldg(p.bias + n + 1); }
                if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                *reinterpret_cast<float2*>(p.out + orow * p.Cout + n) = make_float2(v0, v1);
            }
        }
    }
}

extern "C" int insmos_conv2d_nhwc_tc(const float* in, int32_t H, int32_t W, int32_t Cin,
                                     const float* weight, int32_t mode, int32_t Cout,
                                     const float* bias, int32_t relu, float* out, void* stream) {
    if (!in || !weight || !out || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || mode < 0 || mode > 2) return INSMOS_ERR_INVALID_ARG;
    if (Cin % BK != 0 || Cout % BN != 0) return INSMOS_ERR_UNSUPPORTED;
    ConvP p{in, weight, bias, out, H, W, Cin, Cout, mode, relu};
    constexpr size_t smem = sizeof(float) * 2 * (2 * BM * A_STRIDE + 2 * BK * B_STRIDE);
    static thread_local insmos_smem_cfg_t configured;
    INSMOS_CHECK_CUDA(insmos_ensure_smem(k_conv_nhwc_tc, smem, configured));
    dim3 grid((unsigned)((H * W + BM - 1) / BM), (unsigned)(Cout / BN), mode == 2 ? 4u : 1u);
    k_conv_nhwc_tc<<<grid, CONV_THREADS, smem, (cudaStream_t)stream>>>(p);
    INSMOS_CHECK_LAUNCH("k_conv_nhwc_tc");
    return INSMOS_OK;
}

// SparseConvTensor.dense() + HeightCompression in NHWC: out[(y*W + x)*(C*D) + c*D + z] = feat[i][c]
// (channel order of `dense().view(N, C*D, H, W)`, height_compression.py:26-30)
__global__ void k_dense_scatter_nhwc(const float* __restrict__ feat, const int32_t* __restrict__ coords, int64_t n, int C,
                                     int D, int H, int W, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * C) return;
    const int64_t row = i / C; const int c = (int)(i - row * C);
    const int32_t* q = coords + row * 4;
    const int z = q[1], y = q[2], x = q[3];
    if ((unsigned)z >= (unsigned)D || (unsigned)y >= (unsigned)H || (unsigned)x >= (unsigned)W) return;
    out[((int64_t)y * W + x) * ((int64_t)C * D) + (int64_t)c * D + z] = feat[i];
}
extern "C" int insmos_dense_scatter_nhwc(const float* feat, const int32_t* coords, int64_t n, int32_t C,
                                         int32_t D, int32_t H, int32_t W, float* out, void* stream) {
    if ((n > 0 && (!feat || !coords)) || !out || C <= 0 || D <= 0 || H <= 0 || W <= 0 || n < 0) return INSMOS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    INSMOS_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)C * D * H * W, st));
    if (n == 0) return INSMOS_OK;
    k_dense_scatter_nhwc<<<(unsigned)ceil_div64(n * C, 256), 256, 0, st>>>(feat, coords, n, C, D, H, W, out);
    INSMOS_CHECK_LAUNCH("k_dense_scatter_nhwc");
    return INSMOS_OK;
}
