// micro-benchmark: cycles per tcgen05.mma kind::tf32 (M=128, K=8) for several N, operand patterns and A sources.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/umma_rate tools/micro/umma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../insmos_b200/csrc/umma.cuh"

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// mode 0: SS, 12 MMAs per "stage" walking 4 K-steps x (lo,hi) images like the conv kernels; 4 stage buffers
// mode 1: SS, every MMA uses the same descriptors
// mode 2: TS, A from TMEM (columns 256..), B walking like mode 0
// mode 3: SS like mode 0 but each of the 3 products goes to its own accumulator
__global__ void __launch_bounds__(128, 1) k_rate(int mode, int N, int nstage, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4 * 65536 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i % (200 * 1024 / 4)] = 0.0f;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;
    long long t0 = 0, t1 = 0;
    if (warp == 1 && lane == 0) {
        const uint32_t idesc = umma_idesc_tf32_m128(N);
        const int b_bytes = N * 128;
        const int stage_bytes = 2 * 16384 + 2 * b_bytes;
        const int nbuf = (200 * 1024) / stage_bytes >= 4 ? 4 : (200 * 1024) / stage_bytes;
        t0 = clock64();
        for (int c = 0; c < nstage; ++c) {
            const uint32_t base = smem_u32(smem + (c % nbuf) * stage_bytes);
            const uint64_t a_hi = umma_desc_sw128(base), a_lo = umma_desc_sw128(base + 16384);
            const uint64_t b_hi = umma_desc_sw128(base + 32768), b_lo = umma_desc_sw128(base + 32768 + b_bytes);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint64_t adv = (uint64_t)(j * 2);
                if (mode == 0) {
                    umma_tf32(tmem_base, a_lo + adv, b_hi + adv, idesc, (c | j) != 0);
                    umma_tf32(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
                    umma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, 1u);
                } else if (mode == 1) {
                    umma_tf32(tmem_base, a_hi, b_hi, idesc, (c | j) != 0);
                    umma_tf32(tmem_base, a_hi, b_hi, idesc, 1u);
                    umma_tf32(tmem_base, a_hi, b_hi, idesc, 1u);
                } else if (mode == 2) {
                    const uint32_t ta = tmem_base + 256 + (uint32_t)((c & 1) * 64 + j * 8);
                    umma_tf32_ts(tmem_base, ta + 32, b_hi + adv, idesc, (c | j) != 0);
                    umma_tf32_ts(tmem_base, ta, b_lo + adv, idesc, 1u);
                    umma_tf32_ts(tmem_base, ta, b_hi + adv, idesc, 1u);
                } else {
                    umma_tf32(tmem_base, a_lo + adv, b_hi + adv, idesc, (c | j) != 0);
                    umma_tf32(tmem_base + N, a_hi + adv, b_lo + adv, idesc, (c | j) != 0);
                    umma_tf32(tmem_base + 2 * N, a_hi + adv, b_hi + adv, idesc, (c | j) != 0);
                }
            }
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    const int smem = 200 * 1024 + 1024;
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int nstage = 200;
    for (int grid : {1, 148})
        for (int mode = 0; mode < 4; ++mode)
            for (int N : {32, 64, 128, 256}) {
                if (mode == 3 && N > 128) continue;

                long long h = 0;
                for (int rep = 0; rep < 2; ++rep) {
                    k_rate<<<grid, 128, smem>>>(mode, N, nstage, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("mode %d N %d: %s\n", mode, N, cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                }
                printf("grid %3d mode %d N %3d: %7.1f cycles per MMA (%.0f per 12-MMA stage), %.1f TFLOP/s tf32 per SM-equivalent x148\n",
                       grid, mode, N, (double)h / (nstage * 12), (double)h / nstage,
                       2.0 * 128 * N * 8 * nstage * 12 / ((double)h / 1.965e9) * 148 / 1e12);
            }
    return 0;
}
