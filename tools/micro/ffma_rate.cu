// micro-benchmark: issue rate of FFMA (3-register form) vs the packed fma.rn.f32x2 of sm_100, per SM sub-partition,
// and of FFMA fed by broadcast LDS.128 operands (the inner loop shape of the exact-fp32 sparse-conv kernel).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_rate ffma_rate.cu && ./ffma_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, const float* w, int iters) {
    __shared__ __align__(16) float sw[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    float x0 = out[threadIdx.x & 7], x1 = out[(threadIdx.x + 1) & 7];
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {                                   // 16 independent FFMA chains, register operands
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x0, x1);
        } else if (MODE == 1) {                            // packed: 8 f32x2 chains
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    unsigned long long av, xv, yv;
                    asm("mov.b64 %0, {%1,%2};" : "=l"(av) : "f"(a[i]), "f"(a[i + 1]));
                    asm("mov.b64 %0, {%1,%2};" : "=l"(xv) : "f"(x0), "f"(x0));
                    asm("mov.b64 %0, {%1,%2};" : "=l"(yv) : "f"(x1), "f"(x1));
                    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(av) : "l"(av), "l"(xv), "l"(yv));
                    asm("mov.b64 {%0,%1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(av));
                }
        } else if (MODE == 2) {                            // conv inner loop: per ci, 2 broadcast LDS.128 + 8 FFMA (x from a register)
#pragma unroll
            for (int ci = 0; ci < 16; ++ci) {
                const float4 w0 = *reinterpret_cast<const float4*>(sw + ((it & 7) * 16 + ci) * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(sw + ((it & 7) * 16 + ci) * 8 + 4);
                const float x = a[8 + (ci & 7)];
                a[0] = fmaf(x, w0.x, a[0]); a[1] = fmaf(x, w0.y, a[1]); a[2] = fmaf(x, w0.z, a[2]); a[3] = fmaf(x, w0.w, a[3]);
                a[4] = fmaf(x, w1.x, a[4]); a[5] = fmaf(x, w1.y, a[5]); a[6] = fmaf(x, w1.z, a[6]); a[7] = fmaf(x, w1.w, a[7]);
            }
        } else {                                           // same with packed f32x2: per ci, 2 LDS.128 + 4 FFMA2
#pragma unroll
            for (int ci = 0; ci < 16; ++ci) {
                const ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(sw + ((it & 7) * 16 + ci) * 8);
                const ulonglong2 w1 = *reinterpret_cast<const ulonglong2*>(sw + ((it & 7) * 16 + ci) * 8 + 4);
                const float x = a[8 + (ci & 7)];
                unsigned long long xv, p0, p1, p2, p3;
                asm("mov.b64 %0, {%1,%2};" : "=l"(xv) : "f"(x), "f"(x));
                asm("mov.b64 %0, {%1,%2};" : "=l"(p0) : "f"(a[0]), "f"(a[1]));
                asm("mov.b64 %0, {%1,%2};" : "=l"(p1) : "f"(a[2]), "f"(a[3]));
                asm("mov.b64 %0, {%1,%2};" : "=l"(p2) : "f"(a[4]), "f"(a[5]));
                asm("mov.b64 %0, {%1,%2};" : "=l"(p3) : "f"(a[6]), "f"(a[7]));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p0) : "l"(xv), "l"(w0.x));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p1) : "l"(xv), "l"(w0.y));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p2) : "l"(xv), "l"(w1.x));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p3) : "l"(xv), "l"(w1.y));
                asm("mov.b64 {%0,%1}, %2;" : "=f"(a[0]), "=f"(a[1]) : "l"(p0));
                asm("mov.b64 {%0,%1}, %2;" : "=f"(a[2]), "=f"(a[3]) : "l"(p1));
                asm("mov.b64 {%0,%1}, %2;" : "=f"(a[4]), "=f"(a[5]) : "l"(p2));
                asm("mov.b64 {%0,%1}, %2;" : "=f"(a[6]), "=f"(a[7]) : "l"(p3));
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long*>(out + 65536)[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_smsp, double fma_per_iter) {
    float *out, *w;
    cudaMalloc(&out, 4 * 65536 + 64);
    cudaMalloc(&w, 4096);
    cudaMemset(out, 0, 4 * 65536 + 64);
    cudaMemset(w, 0, 4096);
    const int iters = 2000, threads = 128 * warps_per_smsp;
    k<MODE><<<148, threads>>>(out, w, 10);
    k<MODE><<<148, threads>>>(out, w, iters);
    cudaDeviceSynchronize();
    long long cyc;
    cudaMemcpy(&cyc, out + 65536, 8, cudaMemcpyDeviceToHost);
    // lane-FMAs per cycle per SM
    const double fma = fma_per_iter * iters * threads;
    printf("%-34s warps/SMSP %d: %8lld cycles, %.1f lane-FMA/clk/SM (%s)\n", name, warps_per_smsp, cyc, fma / cyc,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(w);
}

int main() {
    for (int wps = 1; wps <= 4; wps *= 2) {
        run<0>("FFMA reg (16 chains)", wps, 128);
        run<1>("FFMA2 f32x2 (8 chains)", wps, 128);
        run<2>("conv loop: 2 LDS.128 + 8 FFMA", wps, 128);
        run<3>("conv loop: 2 LDS.128 + 4 FFMA2", wps, 128);
    }
    return 0;
}
