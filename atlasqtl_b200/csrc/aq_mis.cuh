// Missing-response sweep: coreDualMisLoop (reference src/coreLoop.cpp:91-138) in sample space.
//
// With missing responses every trait has its own Gram matrix X' diag(mis_k) X (the reference's cp_X - cp_X_rm[[k]],
// R/atlasqtl_global_local_core.R:25-32), so the SNP blocking of the main kernel -- one Gram band shared by all traits --
// does not apply.  The path stays exact and simple instead: ONE WARP PER TRAIT keeps that trait's masked residual
// r_k = mis_k o (y_k - X beta_k) in registers (sample i = lane + 32 m), walks the SNPs in sweep order, and for each
//   s = x_j' r_k + beta_jk X_norm_sq(j,k)        (warp reduction; r is zero in the missing rows)
//   mu, gam, beta  as src/coreLoop.cpp:125-131  with  sig2_beta_vb(j,k) = 1 / (c (X_norm_sq(j,k) + sig2_inv) tau_k)
//   r_k -= delta (mis_k o x_j)                    (predicated on the lane's mask bits)
// The p x q scalars of 32 consecutive SNPs are fetched at once, one SNP per lane (everything that does not depend on
// s -- the divide, the log, the annealed offsets -- is evaluated lane-parallel there), and handed to the serial step by
// shuffles; results and the per-trait running sums are likewise kept one SNP per lane.  X columns stream through
// L1 / L2 (all warps walk them in the same order).  fp64 FMA bound at ~2n FMA per update; n <= 2048.
#pragma once
#include "aq_common.cuh"

namespace aq {

struct MisParams {
    const double* xraw;     // [p][n]
    const int* order;       // [p_pad], -1 padding
    const unsigned long long* mask;  // [q_pad][32]: bit m of word (k, lane) = sample lane + 32 m of trait k is observed
    int n, p_pad, q, q_pad, ld_resid;
    double* resid;          // [q_pad][ld_resid]
    double* gam;            // [p_pad][q_pad]
    double* mu;
    const double* dtab;
    const double* wtab;
    const double* i0tab;
    double* xnsq;           // [p_pad][q_pad]  X_norm_sq = crossprod(X^2, mis_pat)  (R/atlasqtl_global_local_core.R:23)
    const double* sig2tab;  // [p_pad][q_pad]  explicit sig2_beta_vb (stateless entry) or NULL: formed from xnsq on the fly
    const double* tau;      // [q_pad]
    const double* log_tau;
    double c, log_sig2_inv, sig2_inv;
    double* out;            // [kMisOutputs][q_pad]
    int mode;               // 0: sweep;  1: residual r = mis o (y - X beta) and sums from the loaded state;  2: X_norm_sq only
};

// rows of MisParams::out
enum {
    kMisGam = 0,      // sum_j gam
    kMisGamMu2,       // sum_j gam mu^2
    kMisS2Gam,        // sum_j sig2_beta_jk gam                (sweep)   | sum_j beta^2 (mode 1)
    kMisXnGamMu2,     // sum_j xn gam mu^2
    kMisXnS2Gam,      // sum_j xn sig2_beta_jk gam             (sweep)   | sum_j xn gam (mode 1)
    kMisXnBeta2,      // sum_j xn beta^2
    kMisRsq,          // |r_k|^2
    kMisZ,            // sum_j gam W + I0                      (sweep)
    kMisGamLogS2,     // sum_j gam log sig2_beta_jk            (sweep)
    kMisOutputs
};

// n <= 512: X columns are consumed in chunks of 8 values per lane (and re-read from L1 for the rank-1 update) so that the
// kernel fits in 128 registers: 16 warps per SM instead of 8 hide the per-SNP latency (reduction, logistic).  Beyond, a
// column is loaded once per SNP and kept in registers (8 warps per SM): the path is then bound by L1 bandwidth, one
// 8n-byte column read per update (the register-blocked remedy needs per-trait Gram corrections: next round).
template <int M>
__global__ void __launch_bounds__(128, (M <= 16) ? 4 : 2) mis_sweep_kernel(const MisParams P) {
    constexpr int kXC = M <= 16 ? (M < 8 ? M : 8) : M;
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // one warp per trait
    if (k >= P.q) return;
    const unsigned long long mbits = P.mask[(size_t)k * 32 + lane];
    double r[M];
    double* rrow = P.resid + (size_t)k * P.ld_resid;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const int i = lane + 32 * m;
        r[m] = (i < P.n && P.mode != 2) ? rrow[i] : 0.0;
    }
    const double tauk = P.tau[k];
    const double cst = -(P.log_tau[k] + P.log_sig2_inv) / 2;   // src/coreLoop.cpp:108
    double acc[kMisOutputs];
#pragma unroll
    for (int i = 0; i < kMisOutputs; ++i) acc[i] = 0.0;

    for (int b0 = 0; b0 < P.p_pad; b0 += 32) {
        // ---- this lane's SNP of the group: everything that does not depend on the running residual
        const int pos = b0 + lane;
        const int jl = pos < P.p_pad ? P.order[pos] : -1;
        const size_t off = (size_t)(jl < 0 ? 0 : jl) * P.q_pad + k;
        double bo = 0.0, ap = 0.0, a = 0.0, bq = 0.0, xn = 0.0, s2 = 1.0, ww = 0.0, ii = 0.0, go = 0.0, mo = 0.0;
        if (jl >= 0 && P.mode != 2) {
            go = P.gam[off];
            mo = P.mu[off];
            bo = go * mo;
            xn = P.xnsq[off];
            if (P.mode == 0) {
                ww = P.wtab[off];
                ii = P.i0tab[off];
                if (P.sig2tab) {                      // sig2_beta_vb(j,k) handed over by the caller (src/coreLoop.cpp:102)
                    s2 = P.sig2tab[off];
                    a = P.c * s2 * tauk;
                } else {
                    a = 1.0 / (xn + P.sig2_inv);      // c sig2_beta tau: mu = a s                        (:125)
                    s2 = a / (P.c * tauk);            // sig2_beta_vb(j,k), update_sig2_beta_vb_ R/update_vb.R:47
                }
                ap = P.c * (P.dtab[off] - log(s2) / 2 + cst);   // :127-129 without the mu^2 term
                bq = P.c * a * a / (2.0 * s2);        // c mu^2 / (2 sig2_beta) = bq s^2
            }
        }
        double gnew = 0.0, mnew = 0.0, xnew = 0.0;
        for (int t = 0; t < 32; ++t) {
            const int j = __shfl_sync(0xffffffffu, jl, t);
            if (j < 0) continue;   // warp-uniform
            const double* x = P.xraw + (size_t)j * P.n;
            if (P.mode == 2) {   // X_norm_sq(j,k) = sum_i x_ij^2 mis_ik
                double q0 = 0.0, q1 = 0.0;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const int i = lane + 32 * m;
                    const double xv = i < P.n ? __ldg(x + i) : 0.0;
                    if ((mbits >> m) & 1ull) {
                        if (m & 1) q1 = fma(xv, xv, q1);
                        else q0 = fma(xv, xv, q0);
                    }
                }
                const double qs = warp_sum(q0 + q1);
                if (lane == t) xnew = qs;
                continue;
            }
            double dlt;
            if (P.mode == 0) {
                double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
                for (int m0 = 0; m0 < M; m0 += kXC) {
                    double xv[kXC];
#pragma unroll
                    for (int e = 0; e < kXC; ++e) {
                        const int i = lane + 32 * (m0 + e);
                        xv[e] = i < P.n ? __ldg(x + i) : 0.0;
                    }
#pragma unroll
                    for (int e = 0; e < kXC; e += 4) {
                        d0 = fma(xv[e], r[m0 + e], d0);
                        if (e + 1 < kXC) d1 = fma(xv[e + 1], r[m0 + e + 1], d1);
                        if (e + 2 < kXC) d2 = fma(xv[e + 2], r[m0 + e + 2], d2);
                        if (e + 3 < kXC) d3 = fma(xv[e + 3], r[m0 + e + 3], d3);
                    }
                }
                const double dot = warp_sum((d0 + d1) + (d2 + d3));
                const double bo_t = __shfl_sync(0xffffffffu, bo, t), xn_t = __shfl_sync(0xffffffffu, xn, t);
                const double a_t = __shfl_sync(0xffffffffu, a, t), bq_t = __shfl_sync(0xffffffffu, bq, t);
                const double ap_t = __shfl_sync(0xffffffffu, ap, t);
                const double s = fma(bo_t, xn_t, dot);                 // :120, :125
                const double m = a_t * s;                              // :125
                const double gm = logistic_neg<>(fma(s * s, -bq_t, ap_t));   // :127-129
                dlt = fma(gm, m, -bo_t);                               // :131-132
                if (lane == t) { gnew = gm; mnew = m; }
            } else {
                dlt = __shfl_sync(0xffffffffu, bo, t);                 // r = mis o (y - X beta): subtract beta_jk x_j
            }
#pragma unroll
            for (int m0 = 0; m0 < M; m0 += kXC) {   // (the column is still in L1)
                double xv[kXC];
#pragma unroll
                for (int e = 0; e < kXC; ++e) {
                    const int i = lane + 32 * (m0 + e);
                    xv[e] = i < P.n ? __ldg(x + i) : 0.0;
                }
#pragma unroll
                for (int e = 0; e < kXC; ++e)
                    if ((mbits >> (m0 + e)) & 1ull) r[m0 + e] = fma(-dlt, xv[e], r[m0 + e]);   // :132
            }
        }
        // ---- one SNP per lane again: results and running sums
        if (jl >= 0) {
            if (P.mode == 2) {
                P.xnsq[off] = xnew;
            } else if (P.mode == 0) {
                P.gam[off] = gnew;
                P.mu[off] = mnew;
                const double bn = gnew * mnew;
                acc[kMisGam] += gnew;
                acc[kMisGamMu2] = fma(bn, mnew, acc[kMisGamMu2]);
                acc[kMisS2Gam] = fma(s2, gnew, acc[kMisS2Gam]);
                acc[kMisXnGamMu2] = fma(xn * bn, mnew, acc[kMisXnGamMu2]);
                acc[kMisXnS2Gam] = fma(xn * s2, gnew, acc[kMisXnS2Gam]);
                acc[kMisXnBeta2] = fma(xn * bn, bn, acc[kMisXnBeta2]);
                acc[kMisZ] += fma(gnew, ww, ii);
                acc[kMisGamLogS2] = fma(gnew, log(s2), acc[kMisGamLogS2]);
            } else {
                acc[kMisGam] += go;
                acc[kMisGamMu2] = fma(bo, mo, acc[kMisGamMu2]);
                acc[kMisS2Gam] = fma(bo, bo, acc[kMisS2Gam]);
                acc[kMisXnGamMu2] = fma(xn * bo, mo, acc[kMisXnGamMu2]);
                acc[kMisXnS2Gam] = fma(xn, go, acc[kMisXnS2Gam]);
                acc[kMisXnBeta2] = fma(xn * bo, bo, acc[kMisXnBeta2]);
            }
        }
    }
    if (P.mode == 2) return;
    double ss = 0.0;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const int i = lane + 32 * m;
        ss = fma(r[m], r[m], ss);
        if (i < P.n) rrow[i] = r[m];
    }
    acc[kMisRsq] = ss;
#pragma unroll
    for (int i = 0; i < kMisOutputs; ++i) {
        const double v = warp_sum(acc[i]);
        if (lane == 0) P.out[(size_t)i * P.q_pad + k] = v;
    }
}

// mis: [q][n] doubles (R's n x q matrix, 1 observed / 0 missing).  Packs the bit masks and zeroes Y where missing
// (Y[is.na(Y)] <- 0, R/atlasqtl_global_local_core.R:22).  One warp per trait.
__global__ void pack_mask_kernel(const double* __restrict__ mis, int n, int q, int ld, unsigned long long* __restrict__ mask,
                                 double* __restrict__ ymat, double* __restrict__ n_obs) {
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= q) return;
    unsigned long long bits = 0ull;
    double cnt = 0.0;
    for (int m = 0; m < 64; ++m) {
        const int i = lane + 32 * m;
        if (i < n) {
            const bool obs = mis[(size_t)k * n + i] != 0.0;
            if (obs) { bits |= 1ull << m; cnt += 1.0; }
            else ymat[(size_t)k * ld + i] = 0.0;
        }
    }
    mask[(size_t)k * 32 + lane] = bits;
    cnt = warp_sum(cnt);
    if (lane == 0) n_obs[k] = cnt;
}

// ---------------------------------------------------------------- tile-structured missing-response path (aq_sweep.cuh, MIS)
// mis: [q][n] doubles as above.  Bit matrix for the MMA warps' accumulator masks: bit i of word i / 64 of row k = sample i of
// trait k is observed (rows of padding traits and bits of padding samples stay 0); zeroes Y where missing; counts.
__global__ void pack_bits_kernel(const double* __restrict__ mis, int n, int q, int mwords, int ld,
                                 unsigned long long* __restrict__ mbits, double* __restrict__ ymat, double* __restrict__ n_obs) {
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= q) return;
    double cnt = 0.0;
    for (int w = lane; w < mwords; w += 32) {
        unsigned long long bits = 0ull;
        for (int b = 0; b < 64; ++b) {
            const int i = w * 64 + b;
            if (i < n) {
                if (mis[(size_t)k * n + i] != 0.0) { bits |= 1ull << b; cnt += 1.0; }
                else ymat[(size_t)k * ld + i] = 0.0;
            }
        }
        mbits[(size_t)k * mwords + w] = bits;
    }
    cnt = warp_sum(cnt);
    if (lane == 0) n_obs[k] = cnt;
}

// Per-trait Gram band of every (SNP block, 16-trait tile):  gk[tile][b][t * 16 + u][kk] = G[t][u] - sum over the MISSING
// samples i of trait kk of x_it x_iu  =  (X' diag(mis_k) X)[t][u]  (the reference's cp_X - cp_X_rm[[k]],
// R/atlasqtl_global_local_core.R:25-32, restricted to the band the blocked sweep needs: u = 0..7 the NEXT block's SNPs,
// u = 8..15 this block's).  G comes from the tile image's band; the missing samples of a trait from its CSR list.
// Its diagonal is X_norm_sq(j, k) = crossprod(X^2, mis_pat) (:23), written out on the way.
// Grid (nb, tiles), 128 threads = one (t, u) pair each, 16 traits in registers.  Once per order, not per sweep.
__global__ void __launch_bounds__(128) gk_build_kernel(const double* __restrict__ tiles, int nb, int nslice, int n_pad_c, int xs,
                                                       size_t tile_stride, const int* __restrict__ mis_off,
                                                       const int* __restrict__ mis_idx, int q, int q_pad,
                                                       double* __restrict__ gk, double* __restrict__ xnsq) {
    const int b = blockIdx.x, tile = blockIdx.y;
    const int t = threadIdx.x >> 4, u = threadIdx.x & 15;
    const int uu = u & 7;
    const bool next = u < 8;
    const bool have = !next || b + 1 < nb;
    const int bu = next ? b + 1 : b;
    const double g0 = tiles[((size_t)b * nslice) * tile_stride + (size_t)kBlk * xs + t * 16 + u];   // (every slice carries the band)
    double acc[16];
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
        const int k = tile * 16 + kk;
        double a = g0;
        if (have && k < q) {
            const int i0 = mis_off[k], i1 = mis_off[k + 1];
            double sub = 0.0;
            for (int ii = i0; ii < i1; ++ii) {
                const int i = mis_idx[ii];
                const int r = i / n_pad_c, li = i - r * n_pad_c;
                const double xt = tiles[((size_t)b * nslice + r) * tile_stride + (size_t)t * xs + swz(li, t)];
                const double xu = tiles[((size_t)bu * nslice + r) * tile_stride + (size_t)uu * xs + swz(li, uu)];
                sub = fma(xt, xu, sub);
            }
            a -= sub;
        }
        acc[kk] = a;
    }
    double* dst = gk + (((size_t)tile * nb + b) * 128 + threadIdx.x) * 16;
#pragma unroll
    for (int kk = 0; kk < 16; kk += 2) {
        double2 v;
        v.x = acc[kk];
        v.y = acc[kk + 1];
        *reinterpret_cast<double2*>(dst + kk) = v;
    }
    if (u == 8 + t) {   // diagonal: X_norm_sq(j, k) for the SNP in slot t
        const int j = reinterpret_cast<const int*>(tiles + ((size_t)b * nslice) * tile_stride + (size_t)kBlk * xs + 128)[t];
        if (j >= 0) {
#pragma unroll
            for (int kk = 0; kk < 16; ++kk)
                if (tile * 16 + kk < q_pad) xnsq[(size_t)j * q_pad + tile * 16 + kk] = acc[kk];
        }
    }
}

// Per-sweep p x q tables of the tile path: a = c sig2_beta tau = 1 / (X_norm_sq + sig2_inv) and log sig2_beta_vb(j, k) with
// sig2_beta_vb = a / (c tau_k) (update_sig2_beta_vb_, R/update_vb.R:47) -- or, when the caller hands sig2_beta_vb over
// (stateless coreDualMisLoop entry), a = c sig2_beta tau from it.  HBM bound, 24 B per pair.
__global__ void mis_prep_kernel(const double* __restrict__ xnsq, const double* __restrict__ sig2tab,
                                const double* __restrict__ tau, int p, int q, int q_pad, double c, double sig2_inv,
                                double* __restrict__ atab, double* __restrict__ ltab) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= q) return;
    const double ctau = c * tau[k];
    for (int j = blockIdx.y; j < p; j += gridDim.y) {
        const size_t off = (size_t)j * q_pad + k;
        double a, s2;
        if (sig2tab) {
            s2 = sig2tab[off];
            a = s2 * ctau;
        } else {
            a = 1.0 / (xnsq[off] + sig2_inv);
            s2 = a / ctau;
        }
        atab[off] = a;
        ltab[off] = log(s2);
    }
}

}  // namespace aq
