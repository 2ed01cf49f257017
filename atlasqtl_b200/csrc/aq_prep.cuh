// Pre-processing on the device: what prepare_data_ does to X and Y before the core sees them
// (R/prepare_atlasqtl.R:57-83; rm_constant_ / rm_collinear_, R/utils.R:276-343).
//   X <- scale(X)                      centre, divide by the n-1 standard deviation      (:57)
//   drop constant columns              scale() turned them into NaN                       (:59-62, R/utils.R:278)
//   drop exact duplicates, keep first  duplicated(mat, MARGIN = 2)                        (:68-69, R/utils.R:305)
//   Y <- scale(Y, center = TRUE, scale = FALSE)   column means over the observed entries  (:83)
// The predictors come either as raw doubles or as packed genotype calls (2 bits per sample, values 0 / 1 / 2), the form
// SNP panels are stored in: at n = 5000, p = 500k the standardised matrix is 20 GB, the calls are 0.6 GB.  Nothing is
// materialised before the set of kept columns is known: a first pass leaves per-column moments and a 128-bit
// fingerprint of the STANDARDISED column; the host groups equal fingerprints, a second pass verifies every suspected
// duplicate value by value, a third writes the kept columns straight into the context's [p][n] matrix.
// All three evaluate the same expression (x - mean) / sd, so "equal" means bitwise equal doubles, as in R.
// One-off O(np) streaming work, bound by HBM (doubles) or by nothing worth naming (packed calls).
#pragma once
#include "aq_common.cuh"

namespace aq {

struct SrcDouble {  // raw predictor matrix, column-major: column j contiguous
    const double* x;
    int n;
    __device__ __forceinline__ double at(int j, int i) const { return x[(size_t)j * n + i]; }
};
struct SrcGeno {  // sample i of column j: bits 2 (i % 4) .. 2 (i % 4) + 1 of byte i / 4; 3 is not a genotype
    const uint8_t* g;
    int bytes_per_col;
    __device__ __forceinline__ int code(int j, int i) const { return (g[(size_t)j * bytes_per_col + (i >> 2)] >> (2 * (i & 3))) & 3; }
    __device__ __forceinline__ double at(int j, int i) const { return (double)code(j, i); }
};

struct ColStats {
    double mean, sd;             // sd = NaN-free: 0 for a constant column
    unsigned long long h1, h2;   // fingerprint of the standardised column (sum over samples of two mixes of (bits, i))
    int bad;                     // NaN / Inf entries (doubles) or invalid calls (code 3)
    int pad;
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long z, unsigned long long a, unsigned long long b) {
    z = (z ^ (z >> 30)) * a;
    z = (z ^ (z >> 27)) * b;
    return z ^ (z >> 31);
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// The value every pass agrees on.  A zero deviation is +0.0 whatever the sign of x - mean would suggest.
__device__ __forceinline__ double standardise(double x, double mean, double sd) { return (x - mean) / sd; }

__device__ __forceinline__ void fingerprint_add(double z, int i, unsigned long long& h1, unsigned long long& h2) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(z) + 0x9e3779b97f4a7c15ULL * (unsigned long long)(i + 1);
    h1 += mix64(b, 0xbf58476d1ce4e5b9ULL, 0x94d049bb133111ebULL);
    h2 += mix64(b ^ 0xd6e8feb86659fd93ULL, 0xff51afd7ed558ccdULL, 0xc4ceb9fe1a85ec53ULL);
}

// ---------------------------------------------------------------- pass 1: moments + fingerprint, one warp per column
// doubles: two-pass mean / sum of squared deviations in a fixed order (lane-strided, then a shuffle tree)
__global__ void __launch_bounds__(256) col_stats_double_kernel(SrcDouble src, int p, ColStats* __restrict__ out) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= p) return;
    const int lane = threadIdx.x & 31, n = src.n;
    double s = 0.0;
    int bad = 0;
    for (int i = lane; i < n; i += 32) {
        const double x = src.at(j, i);
        bad += !isfinite(x);
        s += x;
    }
    const double mean = warp_sum(s) / n;
    double ss = 0.0;
    for (int i = lane; i < n; i += 32) {
        const double d = src.at(j, i) - mean;
        ss = fma(d, d, ss);
    }
    const double sd = sqrt(warp_sum(ss) / (n - 1));
    unsigned long long h1 = 0, h2 = 0;
    if (sd > 0.0)
        for (int i = lane; i < n; i += 32) fingerprint_add(standardise(src.at(j, i), mean, sd), i, h1, h2);
    h1 = warp_sum_u64(h1);
    h2 = warp_sum_u64(h2);
    bad = warp_sum_int(bad);
    if (lane == 0) out[j] = ColStats{mean, sd > 0.0 ? sd : 0.0, h1, h2, bad, 0};
}
// genotype calls: the moments follow exactly from the call counts (integers), a standardised column takes three values
__global__ void __launch_bounds__(256) col_stats_geno_kernel(SrcGeno src, int n, int p, ColStats* __restrict__ out) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= p) return;
    const int lane = threadIdx.x & 31;
    int c1 = 0, c2 = 0, bad = 0;
    for (int i = lane; i < n; i += 32) {
        const int g = src.code(j, i);
        c1 += g == 1;
        c2 += g == 2;
        bad += g == 3;
    }
    c1 = warp_sum_int(c1);
    c2 = warp_sum_int(c2);
    bad = warp_sum_int(bad);
    const long long S = (long long)c1 + 2LL * c2, S2 = (long long)c1 + 4LL * c2;
    const double mean = (double)S / n;
    const double ss = (double)((long long)n * S2 - S * S) / n;  // sum (x - mean)^2 = (n sum x^2 - (sum x)^2) / n, exact numerator
    const double sd = sqrt(ss / (n - 1));
    unsigned long long h1 = 0, h2 = 0;
    if (sd > 0.0) {
        const double z[3] = {standardise(0.0, mean, sd), standardise(1.0, mean, sd), standardise(2.0, mean, sd)};
        for (int i = lane; i < n; i += 32) {
            const int g = src.code(j, i);
            fingerprint_add(z[g < 3 ? g : 0], i, h1, h2);
        }
    }
    h1 = warp_sum_u64(h1);
    h2 = warp_sum_u64(h2);
    if (lane == 0) out[j] = ColStats{mean, sd > 0.0 ? sd : 0.0, h1, h2, bad, 0};
}

// ---------------------------------------------------------------- pass 2: suspected duplicates, value by value
// pairs[2 m] = later column, pairs[2 m + 1] = the first column of its fingerprint group; differs[m] = 1 for unequal pairs
template <class Src>
__global__ void __launch_bounds__(256) verify_dups_kernel(Src src, int n, const ColStats* __restrict__ st,
                                                           const int* __restrict__ pairs, int npairs, int* __restrict__ differs) {
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= npairs) return;
    const int lane = threadIdx.x & 31;
    const int a = pairs[2 * m], b = pairs[2 * m + 1];
    const ColStats sa = st[a], sb = st[b];
    int diff = 0;
    for (int i = lane; i < n; i += 32) {
        const double za = standardise(src.at(a, i), sa.mean, sa.sd), zb = standardise(src.at(b, i), sb.mean, sb.sd);
        diff += __double_as_longlong(za) != __double_as_longlong(zb);
    }
    diff = warp_sum_int(diff);
    if (lane == 0) differs[m] = diff != 0;
}

// ---------------------------------------------------------------- pass 3: kept columns -> xraw [p_kept][n]
template <class Src>
__global__ void __launch_bounds__(256) materialise_kernel(Src src, int n, const ColStats* __restrict__ st,
                                                           const int* __restrict__ kept, int p_kept, double* __restrict__ xraw) {
    const int jj = blockIdx.x;
    if (jj >= p_kept) return;
    const int j = kept[jj];
    const ColStats s = st[j];
    for (int i = threadIdx.x; i < n; i += blockDim.x) xraw[(size_t)jj * n + i] = standardise(src.at(j, i), s.mean, s.sd);
}

// ---------------------------------------------------------------- Y: centre every trait over its observed entries
// ymat rows [q][ld] hold the raw responses (NaN = missing); afterwards the centred values, 0 in the missing positions
// (the reference zeroes them before the sweep, R/atlasqtl_global_local_core.R:22).  n_mis[k] = number of NaNs of trait k.
__global__ void __launch_bounds__(256) center_y_kernel(double* __restrict__ ymat, int n, int q, int ld, int* __restrict__ n_mis) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= q) return;
    const int lane = threadIdx.x & 31;
    double* y = ymat + (size_t)k * ld;
    double s = 0.0;
    int nm = 0;
    for (int i = lane; i < n; i += 32) {
        const double v = y[i];
        if (isnan(v)) ++nm;
        else s += v;
    }
    nm = warp_sum_int(nm);
    const double mean = warp_sum(s) / (n - nm);
    for (int i = lane; i < n; i += 32) {
        const double v = y[i];
        y[i] = isnan(v) ? 0.0 : v - mean;
    }
    if (lane == 0) n_mis[k] = nm;
}

}  // namespace aq
