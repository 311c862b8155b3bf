// MUFU throughput on one SM: N dependent-free MUFU ops per thread, 1024 threads per CTA, one CTA per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probes/_build/mufu_probe tools/probes/mufu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __device__ __forceinline__ float op(float x) {
    float r;
    if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    if (OP == 2) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    if (OP == 3) asm volatile("fma.rn.f32 %0, %1, %1, %1;" : "=f"(r) : "f"(x));
    if (OP == 4) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    if (OP == 5) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
template <int OP> __global__ void k(float* out, long long* cyc, int iters) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = op<OP>(a[i]);
    __syncthreads();
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const char* names[] = {"tanh.approx", "rcp.approx", "ex2.approx", "fma", "rsqrt.approx", "lg2.approx"};
    for (int threads : {256, 512, 1024}) for (int o = 0; o < 6; ++o) {
        const int iters = 2000;
        for (int rep = 0; rep < 2; ++rep) {
            if (o == 0) k<0><<<148, threads>>>(out, cyc, iters); if (o == 1) k<1><<<148, threads>>>(out, cyc, iters);
            if (o == 2) k<2><<<148, threads>>>(out, cyc, iters); if (o == 3) k<3><<<148, threads>>>(out, cyc, iters);
            if (o == 4) k<4><<<148, threads>>>(out, cyc, iters); if (o == 5) k<5><<<148, threads>>>(out, cyc, iters);
        }
        long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        printf("%-14s threads %4d: %.2f ops / clk / SM\n", names[o], threads, (double)threads * 8 * iters / h[0]);
    }
    return cudaGetLastError() != cudaSuccess;
}
