// Microbenchmark: DMMA throughput when only SM sub-partitions 0-2 issue (9 of 12 warps), vs all 12 warps.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void __launch_bounds__(384, 1) k(double* out, int iters, int skip3, double a, double b) {
    const int wid = threadIdx.x >> 5;
    if (skip3 && (wid & 3) == 3) return;
    double d[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { d[i][0] = i; d[i][1] = threadIdx.x; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) mma884(d[i][0], d[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += d[i][0] + d[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC> void run(double* out, int skip3) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    k<NACC><<<148, 384>>>(out, iters, skip3, 0.999, 1e-3); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<NACC><<<148, 384>>>(out, iters, skip3, 0.999, 1e-3); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int warps = skip3 ? 9 : 12;
    printf("NACC %2d  %s warps: %.2f TFLOP/s\n", NACC, skip3 ? "9 (SMSP 0-2)" : "12 (all)", 2.0 * 256 * NACC * iters * warps * 148 / ms / 1e9);
}
int main() {
    double* out; cudaMalloc(&out, 8 * 148 * 384);
    run<2>(out, 1); run<4>(out, 1); run<8>(out, 1); run<28>(out, 1);
    run<2>(out, 0); run<4>(out, 0); run<28>(out, 0);
    return 0;
}
