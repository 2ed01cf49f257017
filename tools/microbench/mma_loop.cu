// Microbenchmark: the sweep kernel's MMA inner loops in isolation (no barriers, no chain): S phase (accumulators as the
// A operand, X from shared memory) and U phase (rank-8 update of the accumulators), 4 warps on each of SMSPs 0-2.
// Reports fp64-pipe cycles per DMMA per SMSP (16 = peak).   usage: mma_loop.bin [variant]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
constexpr int MT = 2, NT = 9, XS = 1016;
// variant bits: 1 = S phase, 2 = U phase, 4 = operands from registers instead of shared memory, 8 = also SMSP 3
template <int V>
__global__ void __launch_bounds__(512, 1) k(double* out, int iters, long long* cyc, const double* gsrc, int tma_gap) {
    extern __shared__ double sm[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8 * XS + 2048; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
    __syncthreads();
    __shared__ unsigned long long spin_bar, done_flag;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&spin_bar))); done_flag = 0; }
    __syncthreads();
    if ((V & 64) && wid == 3) {   // TMA traffic: one 66 KB bulk copy per ~`tma_gap` cycles into the upper smem region
        if (lane == 0) {
            __shared__ unsigned long long tbar;
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&tbar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const unsigned bytes = (8 * XS + 144) * 8;
            double* dst = sm + 8 * XS + 2048;
            unsigned phase = 0;
            while (*((volatile unsigned long long*)&done_flag) == 0) {
                const long long t0 = clock64();
                asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(&tbar)), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(gsrc + (size_t)(blockIdx.x % 64) * 16384), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(&tbar)) : "memory");
                unsigned ok = 0;
                while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(&tbar)), "r"(phase) : "memory");
                phase ^= 1;
                while (clock64() - t0 < tma_gap) {}
            }
        }
        return;
    }
    if (!(V & 8) && (wid & 3) == 3) {
        if (V & 32) {  // spinners: wait on a barrier that completes only when the MMA warps are done
            unsigned ok = 0;
            while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(&spin_bar)) : "memory");
        }
        return;
    }
    const int g = lane >> 2, l = lane & 3;
    const int ws = wid - (wid >> 2);
    const int i0 = (ws % 14) * NT * 8;
    const int offS = g * XS + ((i0 + 2 * l) ^ ((g & 2) << 1));
    const int offU0 = l * XS + ((i0 + g) ^ ((l & 2) << 1));
    const int offU1 = (l + 4) * XS + ((i0 + g) ^ ((l & 2) << 1));
    double acc[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { acc[mt][nt][0] = mt + nt; acc[mt][nt][1] = lane; }
    double* sp = sm + 8 * XS + wid * 64;
    __shared__ unsigned long long bar[2];   // bar[0]: completed once (parity-0 waits succeed at once); bar[1]: sink for arrives
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"((unsigned)__cvta_generic_to_shared(&bar[1])));
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(&bar[0])) : "memory");
    }
    asm volatile("bar.sync 1, 384;" ::: "memory");
    auto fake_wait = [&]() {
        unsigned ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(&bar[0])) : "memory");
    };
    auto fake_arrive = [&]() {
        __syncwarp();
        if (lane == 0) asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(&bar[1])) : "memory");
    };
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (V & 1) {
            if (V & 16) fake_wait();
            double sa[MT][2][2];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) sa[mt][0][0] = sa[mt][0][1] = sa[mt][1][0] = sa[mt][1][1] = 0.0;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                double2 xb;
                if (V & 4) { xb.x = 1e-3 * nt; xb.y = 2e-3; } else xb = *reinterpret_cast<const double2*>(sm + offS + nt * 8);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) dmma(sa[mt][0][0], sa[mt][0][1], acc[mt][nt][0], xb.x);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) dmma(sa[mt][1][0], sa[mt][1][1], acc[mt][nt][1], xb.y);
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                double2 v;
                v.x = sa[mt][0][0] + sa[mt][1][0];
                v.y = sa[mt][0][1] + sa[mt][1][1];
                *reinterpret_cast<double2*>(sp + (mt * 8 + g) * 2 + 0) = v;   // (layout irrelevant here)
            }
            if (V & 16) fake_arrive();
        }
        if (V & 2) {
            if (V & 16) fake_wait();
            double nd[MT][2];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) { nd[mt][0] = sm[8 * XS + 1024 + (mt * 8 + g) * 8 + l]; nd[mt][1] = sm[8 * XS + 1024 + (mt * 8 + g) * 8 + l + 4]; }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                double x0, x1;
                if (V & 4) { x0 = 1e-3; x1 = 1e-4 * nt; } else { x0 = sm[offU0 + nt * 8]; x1 = sm[offU1 + nt * 8]; }
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) dmma(acc[mt][nt][0], acc[mt][nt][1], nd[mt][0], x0);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) dmma(acc[mt][nt][0], acc[mt][nt][1], nd[mt][1], x1);
            }
            if (V & 16) fake_arrive();
        }
    }
    const long long t1 = clock64();
    if (V & 64) { asm volatile("bar.sync 2, 384;" ::: "memory"); if (threadIdx.x == 0) *((volatile unsigned long long*)&done_flag) = 1; }
    if (V & 32) { asm volatile("bar.sync 2, 384;" ::: "memory"); if (threadIdx.x == 0) asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(&spin_bar)) : "memory"); }
    double s = 0;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) s += acc[mt][nt][0] + acc[mt][nt][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int V> void run(double* out, long long* cyc) {
    const int iters = 2000;
    const size_t smem = (8 * XS + 2048 + ((V & 64) ? 8 * XS + 144 : 0)) * sizeof(double);
    static double* gsrc = nullptr;
    if (!gsrc) { cudaMalloc(&gsrc, 64 * 16384 * 8 + (8 * XS + 144) * 8); cudaMemset(gsrc, 0, 64 * 16384 * 8 + (8 * XS + 144) * 8); }
    cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<V><<<148, 512, smem>>>(out, iters, cyc, gsrc, 4600); cudaDeviceSynchronize();
    k<V><<<148, 512, smem>>>(out, iters, cyc, gsrc, 4600); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const int phases = ((V & 1) ? 1 : 0) + ((V & 2) ? 1 : 0);
    const double dmma_per_smsp = 4.0 * MT * NT * 2 * phases * iters;   // 4 warps per SMSP
    printf("variant %2d (%s%s%s%s%s%s%s): %.2f cycles per DMMA per SMSP (%s)\n", V, (V & 1) ? "S " : "", (V & 2) ? "U " : "",
           (V & 4) ? "reg-operands " : "smem-operands ", (V & 16) ? "barriers " : "", (V & 32) ? "4-spinners " : "", (V & 64) ? "TMA-traffic " : "", (V & 8) ? "4 SMSPs" : "3 SMSPs", (double)h / dmma_per_smsp, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 148 * 512); cudaMalloc(&cyc, 64);
    run<3>(out, cyc); run<2>(out, cyc); run<7>(out, cyc); run<19>(out, cyc); run<83>(out, cyc); run<67>(out, cyc);
    return 0;
}
