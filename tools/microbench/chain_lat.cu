// Microbenchmark: dependent-chain latency of the logistic variants used by the sweep's chain warp.
#include <cstdio>
#include "../../atlasqtl_b200/csrc/aq_common.cuh"
__global__ void k(double* out, long long* cyc, int iters, double x0) {
    double s = x0 + threadIdx.x * 1e-9; long long t0, t1;
    t0 = clock64(); for (int i = 0; i < iters; i++) s = 1.0 / (1.0 + exp(s)); t1 = clock64(); if (!threadIdx.x) cyc[0] = t1 - t0;
    double a = x0;
    t0 = clock64(); for (int i = 0; i < iters; i++) a = aq::logistic_neg(a); t1 = clock64(); if (!threadIdx.x) cyc[1] = t1 - t0;
    // accuracy of the fast logistic against libm over [-40, 40] (max relative error, reported in units of 1e-16)
    double b = 0.0;
    for (int i = 0; i < iters; i++) { const double x = -40.0 + 80.0 * (i + threadIdx.x / 32.0) / iters; const double r = 1.0 / (1.0 + exp(x)); b = fmax(b, fabs(aq::logistic_neg(x) - r) / r); }
    b = fmax(b, __shfl_xor_sync(0xffffffffu, b, 16)); b = fmax(b, __shfl_xor_sync(0xffffffffu, b, 8)); b = fmax(b, __shfl_xor_sync(0xffffffffu, b, 4));
    b = fmax(b, __shfl_xor_sync(0xffffffffu, b, 2)); b = fmax(b, __shfl_xor_sync(0xffffffffu, b, 1));
    if (!threadIdx.x) { cyc[2] = (long long)(b * 1e16 * iters); cyc[3] = 0; }
    double c = 0.0;
    double d = x0 + 1.0;
    t0 = clock64(); for (int i = 0; i < iters; i++) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d)); d = y + 1.0; } t1 = clock64(); if (!threadIdx.x) cyc[4] = t1 - t0;
    double e = x0;
    t0 = clock64(); for (int i = 0; i < iters; i++) e = exp(-e); t1 = clock64(); if (!threadIdx.x) cyc[5] = t1 - t0;
    out[threadIdx.x] = s + a + b + c + d + e;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 32); cudaMalloc(&cyc, 8 * 8);
    int it = 2000;
    k<<<1, 32>>>(out, cyc, it, 0.3); cudaDeviceSynchronize(); k<<<1, 32>>>(out, cyc, it, 0.3); cudaDeviceSynchronize();
    long long h[6]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const char* nm[6] = {"1/(1+exp(x)) libm", "logistic_neg (fast)", "logistic_neg max rel err /1e-16", "-", "rcp.approx.ftz.f64(+add)", "exp libm"};
    for (int i = 0; i < 6; i++) printf("%-28s %.1f cyc/iter\n", nm[i], (double)h[i] / it);
    return 0;
}
