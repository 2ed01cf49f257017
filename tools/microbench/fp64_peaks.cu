// Microbenchmark: fp64 pipe peaks on B200 (sm_100a) -- DFMA vs DMMA (mma.sync f64 shapes),
// plus dependent-chain latencies of the fp64 ops the in-block Gauss-Seidel chain uses.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void k_dfma(double* out, int iters, double a, double b) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void mma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double* d, double a0, double a1, double b0) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a0), "d"(a1), "d"(b0));
}
__device__ __forceinline__ void mma1688(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void k_mma884(double* out, int iters, double a, double b) {
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
template <int NACC, int SHAPE>
__global__ void k_mma16(double* out, int iters, double a, double b) {
    double d[NACC][4];
    double av[8], bv[4];
#pragma unroll
    for (int i = 0; i < 8; i++) av[i] = a + i;
#pragma unroll
    for (int i = 0; i < 4; i++) bv[i] = b + i;
#pragma unroll
    for (int i = 0; i < NACC; i++) { d[i][0] = i; d[i][1] = threadIdx.x; d[i][2] = 1; d[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (SHAPE == 4) mma1684(d[i], av[0], av[1], bv[0]);
            if (SHAPE == 8) mma1688(d[i], av, bv);
            if (SHAPE == 16) mma16816(d[i], av, bv);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// latency chains: one warp, one block
__global__ void k_lat(double* out, long long* cyc, int iters, double x0) {
    double x = x0 + threadIdx.x * 1e-9;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
    for (int i = 0; i < iters; i++) x = fma(x, 0.999999, 1e-7);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // exp chain
    double y = x * 1e-3;
    t0 = clock64();
    for (int i = 0; i < iters; i++) y = exp(-y) ;
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // div chain
    double z = x;
    t0 = clock64();
    for (int i = 0; i < iters; i++) z = 1.0 / (1.0 + z);
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // log chain
    double w = x + 2.0;
    t0 = clock64();
    for (int i = 0; i < iters; i++) w = log(w) + 2.0;
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // sigmoid as 1/(1+exp(x)) chain
    double s = x * 1e-3;
    t0 = clock64();
    for (int i = 0; i < iters; i++) s = 1.0 / (1.0 + exp(s));
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // dependent mma m8n8k4 chain
    double d0 = x, d1 = y;
    t0 = clock64();
    for (int i = 0; i < iters; i++) mma884(d0, d1, 0.5, 0.25);
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // rcp via fast path
    double r = x + 1.5;
    t0 = clock64();
    for (int i = 0; i < iters; i++) r = __drcp_rn(r) + 1.5;
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // shfl double chain
    double sh = x;
    t0 = clock64();
    for (int i = 0; i < iters; i++) sh = __shfl_xor_sync(0xffffffffu, sh, 1) + 1.0;
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
    out[threadIdx.x] = x + y + z + w + s + d0 + d1 + r + sh;
}

// best-of-`reps` CUDA-event time of one launch; < 0 if the launch itself fails (e.g. too many registers for the block size)
template <typename F>
float timeit(F f, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f();
    if (cudaGetLastError() != cudaSuccess) return -1.f;
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", prop.name, sms, prop.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    long long* cyc; CK(cudaMalloc(&cyc, 64 * sizeof(long long)));
    int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        int threads = warps * 32; int blocks = sms * (warps <= 16 ? 2 : 1);
        float ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters, 0.999, 1e-3); });
        double flops = 2.0 * 16 * iters * (double)threads * blocks;
        if (ms > 0) printf("DFMA   blocks %d x %d thr: %.2f TFLOP/s\n", blocks, threads, flops / ms / 1e9);
    }
    for (int warps : {4, 8, 16, 32}) {
        int threads = warps * 32; int blocks = sms * (warps <= 16 ? 2 : 1);
        float ms = timeit([&] { k_mma884<8><<<blocks, threads>>>(out, iters, 0.999, 1e-3); });
        double flops = 2.0 * 256 * 8 * iters * (double)warps * blocks;
        if (ms > 0) printf("DMMA m8n8k4   x8acc blocks %d x %d thr: %.2f TFLOP/s\n", blocks, threads, flops / ms / 1e9);
        else printf("DMMA m8n8k4   x8acc blocks %d x %d thr: not launchable (registers)\n", blocks, threads);
        ms = timeit([&] { k_mma884<2><<<blocks, threads>>>(out, iters, 0.999, 1e-3); });
        flops = 2.0 * 256 * 2 * iters * (double)warps * blocks;
        if (ms > 0) printf("DMMA m8n8k4   x2acc blocks %d x %d thr: %.2f TFLOP/s\n", blocks, threads, flops / ms / 1e9);
        else printf("DMMA m8n8k4   x2acc blocks %d x %d thr: not launchable (registers)\n", blocks, threads);
        ms = timeit([&] { k_mma16<8, 4><<<blocks, threads>>>(out, iters, 0.999, 1e-3); });
        flops = 2.0 * 512 * 8 * iters * (double)warps * blocks;
        if (ms > 0) printf("DMMA m16n8k4  x8acc blocks %d x %d thr: %.2f TFLOP/s\n", blocks, threads, flops / ms / 1e9);
        else printf("DMMA m16n8k4  x8acc blocks %d x %d thr: not launchable (registers)\n", blocks, threads);
        ms = timeit([&] { k_mma16<8, 8><<<blocks, threads>>>(out, iters, 0.999, 1e-3); });
        flops = 2.0 * 1024 * 8 * iters * (double)warps * blocks;
        if (ms > 0) printf("DMMA m16n8k8  x8acc blocks %d x %d thr: %.2f TFLOP/s\n", blocks, threads, flops / ms / 1e9);
        else printf("DMMA m16n8k8  x8acc blocks %d x %d thr: not launchable (registers)\n", blocks, threads);
        ms = timeit([&] { k_mma16<8, 16><<<blocks, threads>>>(out, iters, 0.999, 1e-3); });
        flops = 2.0 * 2048 * 8 * iters * (double)warps * blocks;
        if (ms > 0) printf("DMMA m16n8k16 x8acc blocks %d x %d thr: %.2f TFLOP/s\n", blocks, threads, flops / ms / 1e9);
        else printf("DMMA m16n8k16 x8acc blocks %d x %d thr: not launchable (registers)\n", blocks, threads);
    }
    int lit = 2000;
    k_lat<<<1, 32>>>(out, cyc, lit, 0.5); CK(cudaDeviceSynchronize());
    k_lat<<<1, 32>>>(out, cyc, lit, 0.5); CK(cudaDeviceSynchronize());
    long long h[8]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    const char* names[8] = {"dfma", "exp", "div 1/(1+z)", "log(+add)", "sigmoid 1/(1+exp)", "dmma m8n8k4", "__drcp_rn(+add)", "shfl64(+add)"};
    for (int i = 0; i < 8; i++) printf("latency %-20s %.1f cyc/iter\n", names[i], (double)h[i] / lit);
    return 0;
}
