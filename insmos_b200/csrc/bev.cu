// Sparse -> dense scatter of the BEV head: SparseConvTensor.dense() + HeightCompression straight
// into the channels-last [H*W, C*D] layout the tensor-core convolutions read (height_compression.py:14-33).  The dense
// convolutions themselves run on tcgen05 (spconv_umma.cu: 3x3 convs through the TMEM-operand kernel; bev_tcgen05.cu: the 2x2
// transposed conv); the round-1 mma.sync implicit-GEMM kernel that used to live here was removed in round 2.
#include "common.cuh"

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
