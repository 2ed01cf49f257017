// Selection sets on the device: {PPI > thres} and {bFDR < thres} (reference R/summarise_output.R:99-106, :207-223)
// without sorting -- or even downloading -- the p x q matrix of posterior inclusion probabilities.
//
// assign_bFDR sorts all PPIs decreasingly and takes the running mean of e = 1 - PPI.  That running mean is
// non-decreasing along the sorted order (every new element is >= all previous ones), so {bFDR < thres} is a PREFIX of
// the sorted order: everything with e <= t1, plus possibly the first few (in column-major index order: R's order() is
// stable) of the elements tied at the next value t2.  The host finds t1 by bisection over the bit pattern of a double
// (ordered like the value for e >= 0), each probe being one streaming pass:
//   ppi_count_sum_kernel   N(t) = #{e <= t}, S(t) = sum{e : e <= t}     (fixed-order reductions: deterministic)
//   ppi_next_kernel        min{e : e > t}
//   ppi_collect_kernel     compaction of the pairs in a value range (or with PPI > thres) into (j, k, PPI) lists
#pragma once
#include "aq_common.cuh"

namespace aq {

constexpr int kSelBlocks = 592;   // 4 x 148 persistent blocks of 256 threads

__device__ __forceinline__ double block_sum_256(double v, double* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < 8; ++w) s += red[w];
    __syncthreads();
    return s;   // valid in thread 0
}

// partial[2 b] = count, partial[2 b + 1] = sum over block b's share; rows j < p, traits k < q of gam [.][q_pad]
__global__ void __launch_bounds__(256) ppi_count_sum_kernel(const double* __restrict__ gam, int p, int q, int q_pad, double t,
                                                            double* __restrict__ partial) {
    __shared__ double red[8];
    double cnt = 0.0, sum = 0.0;
    for (int j = blockIdx.x; j < p; j += gridDim.x) {
        const double* row = gam + (size_t)j * q_pad;
        for (int k = threadIdx.x; k < q; k += 256) {
            const double e = 1.0 - row[k];
            if (e <= t) { cnt += 1.0; sum += e; }
        }
    }
    const double c = block_sum_256(cnt, red);
    const double s = block_sum_256(sum, red);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = c; partial[2 * blockIdx.x + 1] = s; }
}

__global__ void __launch_bounds__(256) ppi_next_kernel(const double* __restrict__ gam, int p, int q, int q_pad, double t,
                                                       double* __restrict__ partial) {
    __shared__ double red[8];
    double mn = __longlong_as_double(0x7ff0000000000000ll);   // +inf
    for (int j = blockIdx.x; j < p; j += gridDim.x) {
        const double* row = gam + (size_t)j * q_pad;
        for (int k = threadIdx.x; k < q; k += 256) {
            const double e = 1.0 - row[k];
            if (e > t) mn = fmin(mn, e);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mn;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) mn = fmin(mn, red[w]);
        partial[blockIdx.x] = mn;
    }
}

// out[0] = sum_b partial[2b], out[1] = sum_b partial[2b+1] (mode 0);  out[0] = min_b partial[b] (mode 1).  One block.
__global__ void ppi_finish_kernel(const double* __restrict__ partial, int nblocks, int mode, double* __restrict__ out) {
    __shared__ double red[8];
    if (mode == 0) {
        double c = 0.0, s = 0.0;
        for (int b = threadIdx.x; b < nblocks; b += 256) { c += partial[2 * b]; s += partial[2 * b + 1]; }
        const double cc = block_sum_256(c, red);
        const double ss = block_sum_256(s, red);
        if (threadIdx.x == 0) { out[0] = cc; out[1] = ss; }
    } else {
        double mn = __longlong_as_double(0x7ff0000000000000ll);
        for (int b = threadIdx.x; b < nblocks; b += 256) mn = fmin(mn, partial[b]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mn;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) mn = fmin(mn, red[w]);
            out[0] = mn;
        }
    }
}

// mode 0: pairs with lo < 1 - gam <= hi;  mode 1: pairs with gam > lo.  Appends (j, k, gam) in arbitrary order;
// *counter ends as the number of matching pairs, of which the first `capacity` were written.
__global__ void __launch_bounds__(256) ppi_collect_kernel(const double* __restrict__ gam, int p, int q, int q_pad, int mode,
                                                          double lo, double hi, long long capacity, int* __restrict__ out_j,
                                                          int* __restrict__ out_k, double* __restrict__ out_g,
                                                          unsigned long long* __restrict__ counter) {
    for (int j = blockIdx.x; j < p; j += gridDim.x) {
        const double* row = gam + (size_t)j * q_pad;
        for (int k0 = 0; k0 < q; k0 += 256) {
            const int k = k0 + threadIdx.x;
            bool hit = false;
            double g = 0.0;
            if (k < q) {
                g = row[k];
                const double e = 1.0 - g;
                hit = mode == 0 ? (e > lo && e <= hi) : (g > lo);
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, hit);
            if (ballot) {   // one atomic per warp with hits
                const int lane = threadIdx.x & 31;
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(ballot));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (hit) {
                    const unsigned long long slot = base + __popc(ballot & ((1u << lane) - 1u));
                    if ((long long)slot < capacity) { out_j[slot] = j; out_k[slot] = k; out_g[slot] = g; }
                }
            }
        }
    }
}

}  // namespace aq
