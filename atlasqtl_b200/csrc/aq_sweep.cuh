// The CAVI Gauss-Seidel sweep kernel (sm_100a): coreDualLoop (reference src/coreLoop.cpp:38-86) in
// sample space, blocked over SNPs, one persistent CTA per trait tile.
//
// Roles inside a CTA (warp-specialised, no __syncthreads in the steady state).  The fp64 tensor op
// (DMMA) and scalar fp64 math share one pipe per SM sub-partition (SMSP = warp id % 4); a chain lane's
// dependent DFMA/DADD sequence queued behind 16-cycle DMMAs runs ~2-3x slower (ncu: stall_math on the
// chain warp, profiles/r1_ncu_sweep_c4_v1.txt).  So the 9 MMA warps live on SMSPs 0-2 (warp id % 4 != 3)
// and the serial chain gets SMSP 3 to itself (warp ids 3, 7, 11):
//   9 "MMA" warps   hold the tile's residual R^T (traits x samples) in REGISTERS as the accumulator
//                   fragments of the rank-8 update  R^T -= Delta^T X_b^T  (DMMA m8n8k4), and reuse the very
//                   same registers as the A operand of  S^T = R^T X_b  -- the accumulator layout
//                   C[m][2l+e] is an A fragment A[m][l] once the contraction index is read as 2l+e, so the
//                   residual never moves: it is loaded once per tile and stored once per tile.
//   1-2 "chain" warps  one lane per trait: resolve the in-block Gauss-Seidel order exactly from the Gram
//                   band (S[t] -= G[t][u] Delta[u]), evaluate mu / gam (annealed logistic) / beta, emit
//                   Delta and the per-trait running sums.
//   1 producer warp one elected lane streams the pre-tiled X blocks (+ Gram band + SNP ids) into a
//                   3-stage shared-memory ring with 1-D bulk copies (TMA engine) on mbarriers.
//
// One-block look-ahead hides the serial chain behind the tensor pipe: S'_{b+1} = X_{b+1}' R_{b-1} is formed
// while the chain of block b runs, and corrected by the cross Gram block, S_{b+1} = S'_{b+1} - G_{b+1,b} Delta_b.
#pragma once
#include "aq_common.cuh"

namespace aq {

struct SweepParams {
    const double* xtiles;   // nb tile images, tile_stride doubles apart
    size_t tile_stride;
    int nb;                 // number of SNP blocks = p_pad / 8
    int ntiles;             // trait tiles = q_pad / kT
    int q;                  // valid traits
    int q_pad;              // leading dimension of the p x q arrays (trait-contiguous)
    int ld_resid;           // leading dimension of resid (samples per trait row) = kNPad
    double* resid;          // [q_pad][ld_resid]
    double* gam;            // [p_pad][q_pad]
    double* mu;             // [p_pad][q_pad]
    const double* dtab;     // [p_pad][q_pad]  log(1-Phi) - log(Phi)
    const double* wtab;     // [p_pad][q_pad]  imr1 - imr0
    const double* i0tab;    // [p_pad][q_pad]  imr0
    const double* tau;      // [q_pad]
    const double* log_tau;  // [q_pad]
    const double* sig2_beta;  // [q_pad]
    double c;
    double log_sig2_inv;
    double* cs_gam;         // [q_pad] outputs
    double* cs_gmu2;
    double* cs_b2;
    double* rsq;
    double* cs_z;
    int mode;               // 0: sweep;  1: build residual (R -= X beta) + sums from the loaded state
};

template <int WS_, int WT_, int MT_, int NT_>
struct SweepCfg {
    static constexpr int WS = WS_;  // MMA warps along samples (split-K of the S GEMM)
    static constexpr int WT = WT_;  // MMA warps along traits
    static constexpr int MT = MT_;  // 8-trait M tiles per MMA warp
    static constexpr int NT = NT_;  // 8-sample N tiles per MMA warp
    static constexpr int kMmaWarps = WS * WT;
    static constexpr int kT = WT * MT * 8;           // traits per tile
    static constexpr int kChainWarps = (kT + 31) / 32;
    static constexpr int kNPad = WS * NT * 8;        // samples, padded
    static constexpr int kXS = kNPad + ((kNPad % 16 == 0) ? 8 : 0);  // tile row stride == 8 (mod 16) doubles
    static constexpr int kThreads = 12 * 32;  // warps 3, 7, 11 (SMSP 3): chain warp(s) + producer
    static constexpr int kStages = 3;
    static constexpr size_t kTileDoubles = (size_t)kBlk * kXS + kTileTail;
    static constexpr int kSps = 10;  // S-partial row stride (doubles): 8 SNP slots + 2 pad => conflict-free 16-byte accesses
    static constexpr size_t kSpartDoubles = (size_t)2 * WS * kT * kSps;
    static constexpr size_t kDbufDoubles = (size_t)2 * kT * kBlk;
    static constexpr size_t kRsqDoubles = (size_t)WS * kT;
    static constexpr size_t kSmemBytes =
        (kStages * kTileDoubles + kSpartDoubles + kDbufDoubles + kRsqDoubles) * sizeof(double) + 16 * sizeof(uint64_t);
    static_assert(kMmaWarps == 9, "9 MMA warps: three per SMSP on SMSPs 0-2");
    static_assert(kChainWarps <= 2, "at most 64 traits per tile");
    static_assert(kXS % 16 == 8, "row stride must be 8 mod 16 doubles");
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::kThreads, 1) sweep_kernel(const SweepParams P) {
    constexpr int WS = Cfg::WS, MT = Cfg::MT, NT = Cfg::NT, kT = Cfg::kT, XS = Cfg::kXS;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);
    double* spart = tiles + kStages * Cfg::kTileDoubles;  // [2][WS][kT][kSps]
    double* dbuf = spart + Cfg::kSpartDoubles;            // [2][kT][kBlk]  (holds -Delta)
    double* rsqs = dbuf + Cfg::kDbufDoubles;              // [WS][kT]
    uint64_t* bars = reinterpret_cast<uint64_t*>(rsqs + Cfg::kRsqDoubles);
    uint64_t* full = bars;             // [kStages]  tile landed (tx bytes)
    uint64_t* empty = bars + kStages;  // [kStages]  8 MMA warps released the tile
    uint64_t* sdone = bars + 2 * kStages;      // [2]  8 MMA warps wrote their S partials
    uint64_t* dready = bars + 2 * kStages + 2; // [2]  chain warps published -Delta

    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // role map: SMSP 3 (wid % 4 == 3) hosts the special warps, SMSPs 0-2 the MMA warps
    const bool is_special = (wid & 3) == 3;
    const int mma_idx = wid - (wid >> 2);   // 0..8 for MMA warps
    const int special_idx = wid >> 2;       // 0..2 for special warps
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], Cfg::kMmaWarps); }
        for (int s = 0; s < 2; ++s) { mbar_init(&sdone[s], Cfg::kMmaWarps); mbar_init(&dready[s], Cfg::kChainWarps); }
        fence_mbar_init();
    }
    __syncthreads();

    const int nb = P.nb;
    const int my_tiles = (P.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t tile_bytes = (uint32_t)(Cfg::kTileDoubles * sizeof(double));

    if (is_special && special_idx == Cfg::kChainWarps) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            const long total = (long)my_tiles * nb;
            for (long it = 0; it < total; ++it) {
                const int stage = (int)(it % kStages);
                const long use = it / kStages;
                if (use > 0) mbar_wait(&empty[stage], (uint32_t)((use - 1) & 1));
                const int b = (int)(it % nb);
                mbar_arrive_expect_tx(&full[stage], tile_bytes);
                bulk_g2s(tiles + stage * Cfg::kTileDoubles, P.xtiles + (size_t)b * P.tile_stride, tile_bytes, &full[stage]);
            }
        }
    } else if (!is_special) {
        // ------------------------------------------------------------------ MMA warps
        const int ws = mma_idx % WS, wt = mma_idx / WS;
        const int mtid = mma_idx * 32 + lane;
        const int g = lane >> 2, l = lane & 3;
        const int i0 = ws * NT * 8;
        const int tr0 = wt * MT * 8;
        // lane-constant shared-memory offsets of the two operand patterns
        const int offS = g * XS + ((i0 + 2 * l) ^ ((g & 2) << 1));         // + nt*8   (16-byte loads)
        const int offU0 = l * XS + ((i0 + g) ^ ((l & 2) << 1));            // ks = 0, + nt*8
        const int offU1 = (l + 4) * XS + ((i0 + g) ^ ((l & 2) << 1));      // ks = 1 (snp l+4 has the same bit 1)
        double acc[MT][NT][2];
        long gb = 0;  // global block counter of this CTA (drives ring stage and barrier parity)
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
            const int k0 = tile * kT;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double2 v = *reinterpret_cast<const double2*>(
                        P.resid + (size_t)(k0 + tr0 + mt * 8 + g) * P.ld_resid + i0 + nt * 8 + 2 * l);
                    acc[mt][nt][0] = v.x;
                    acc[mt][nt][1] = v.y;
                }
            auto s_phase = [&](long gbi) {
                const int stage = (int)(gbi % kStages);
                mbar_wait(&full[stage], (uint32_t)((gbi / kStages) & 1));
                const double* xt = tiles + stage * Cfg::kTileDoubles;
                double sa[MT][2][2];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) sa[mt][0][0] = sa[mt][0][1] = sa[mt][1][0] = sa[mt][1][1] = 0.0;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double2 xb = *reinterpret_cast<const double2*>(xt + offS + nt * 8);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        dmma(sa[mt][0][0], sa[mt][0][1], acc[mt][nt][0], xb.x);
                        dmma(sa[mt][1][0], sa[mt][1][1], acc[mt][nt][1], xb.y);
                    }
                }
                double* sp = spart + ((size_t)(gbi & 1) * WS + ws) * kT * Cfg::kSps;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    double2 v;  // C fragment: S^T[trait g + 8 mt][snp 2l, 2l + 1]
                    v.x = sa[mt][0][0] + sa[mt][1][0];
                    v.y = sa[mt][0][1] + sa[mt][1][1];
                    *reinterpret_cast<double2*>(sp + (tr0 + mt * 8 + g) * Cfg::kSps + 2 * l) = v;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sdone[gbi & 1]);
            };
            if (P.mode == 0) s_phase(gb);
            for (int b = 0; b < nb; ++b, ++gb) {
                if (P.mode == 0 && b + 1 < nb) s_phase(gb + 1);
                // ---- rank-8 update with -Delta_b
                const int stage = (int)(gb % kStages);
                if (P.mode != 0) mbar_wait(&full[stage], (uint32_t)((gb / kStages) & 1));  // sweep mode: S phase waited
                mbar_wait(&dready[gb & 1], (uint32_t)((gb >> 1) & 1));
                const double* xt = tiles + stage * Cfg::kTileDoubles;
                const double* db = dbuf + (size_t)(gb & 1) * kT * kBlk;
                double nd[MT][2];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    nd[mt][0] = db[(tr0 + mt * 8 + g) * kBlk + l];
                    nd[mt][1] = db[(tr0 + mt * 8 + g) * kBlk + l + 4];
                }
                if (P.mode != 0) {
                    // no S phase paces the chain in this mode: tell it that -Delta buffer (gb & 1) has been consumed,
                    // otherwise it could run two blocks ahead, overwrite the buffer and alias the barrier phase
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sdone[gb & 1]);
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double x0 = xt[offU0 + nt * 8];
                    const double x1 = xt[offU1 + nt * 8];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        dmma(acc[mt][nt][0], acc[mt][nt][1], nd[mt][0], x0);
                        dmma(acc[mt][nt][0], acc[mt][nt][1], nd[mt][1], x1);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
            }
            // ---- tile epilogue: store the residual, per-trait squared norms
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                double ss = 0.0;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    double2 v;
                    v.x = acc[mt][nt][0];
                    v.y = acc[mt][nt][1];
                    ss = fma(v.x, v.x, ss);
                    ss = fma(v.y, v.y, ss);
                    *reinterpret_cast<double2*>(P.resid + (size_t)(k0 + tr0 + mt * 8 + g) * P.ld_resid + i0 + nt * 8 + 2 * l) = v;
                }
                ss += __shfl_xor_sync(0xffffffffu, ss, 1);
                ss += __shfl_xor_sync(0xffffffffu, ss, 2);
                if (l == 0) rsqs[ws * kT + tr0 + mt * 8 + g] = ss;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(Cfg::kMmaWarps * 32) : "memory");
            if (mtid < kT) {
                double ss = 0.0;
#pragma unroll
                for (int w2 = 0; w2 < WS; ++w2) ss += rsqs[w2 * kT + mtid];
                if (k0 + mtid < P.q) P.rsq[k0 + mtid] = ss;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(Cfg::kMmaWarps * 32) : "memory");
        }
    } else if (special_idx < Cfg::kChainWarps) {
        // ------------------------------------------------------------------ chain warps (one lane per trait)
        const int cw = special_idx;
        const int tl = cw * 32 + lane;
        const bool active = tl < kT;
        const int tls = active ? tl : 0;  // inactive lanes shadow trait 0 without side effects
        long gb = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
            const int k = tile * kT + tls;
            const bool valid = active && k < P.q;
            const double sig2 = P.sig2_beta[k], tauk = P.tau[k];
            const double a = P.c * sig2 * tauk;                                   // src/coreLoop.cpp:73
            const double hinv = 1.0 / (2.0 * sig2);                               // :76
            const double cst = -(P.log_tau[k] + P.log_sig2_inv + log(sig2)) / 2;  // :56
            double sg = 0.0, sgm2 = 0.0, sb2 = 0.0, sz = 0.0;
            double dprev[kBlk];
#pragma unroll
            for (int t = 0; t < kBlk; ++t) dprev[t] = 0.0;
            for (int b = 0; b < nb; ++b, ++gb) {
                const int stage = (int)(gb % kStages);
                mbar_wait(&full[stage], (uint32_t)((gb / kStages) & 1));
                const double* xt = tiles + stage * Cfg::kTileDoubles;
                const double* gband = xt + kBlk * XS;
                const int* ids = reinterpret_cast<const int*>(gband + 128);
                double go[kBlk], mo[kBlk], dd[kBlk], ww[kBlk], ii[kBlk];
                int id[kBlk];
#pragma unroll
                for (int t = 0; t < kBlk; ++t) {
                    id[t] = ids[t];
                    const size_t off = (size_t)(id[t] < 0 ? 0 : id[t]) * P.q_pad + k;
                    go[t] = P.gam[off];
                    mo[t] = P.mu[off];
                    if (P.mode == 0) {
                        dd[t] = P.dtab[off];
                        ww[t] = P.wtab[off];
                        ii[t] = P.i0tab[off];
                    } else {
                        dd[t] = ww[t] = ii[t] = 0.0;
                    }
                }
                double s[kBlk], dl[kBlk];
                if (P.mode == 0) {
                    mbar_wait(&sdone[gb & 1], (uint32_t)((gb >> 1) & 1));
                    const double* sp = spart + (size_t)(gb & 1) * WS * kT * Cfg::kSps + tls * Cfg::kSps;
#pragma unroll
                    for (int t = 0; t < kBlk; ++t) s[t] = 0.0;
#pragma unroll
                    for (int w2 = 0; w2 < WS; ++w2) {
#pragma unroll
                        for (int t = 0; t < kBlk; t += 2) {
                            const double2 v = *reinterpret_cast<const double2*>(sp + w2 * kT * Cfg::kSps + t);
                            s[t] += v.x;
                            s[t + 1] += v.y;
                        }
                    }
                    // look-ahead correction: S was formed before the previous block's update was applied
                    if (b > 0) {
#pragma unroll
                        for (int t = 0; t < kBlk; ++t)
#pragma unroll
                            for (int u = 0; u < kBlk; ++u) s[t] = fma(-gband[t * 16 + u], dprev[u], s[t]);
                    }
#pragma unroll
                    for (int t = 0; t < kBlk; ++t) {
                        const double bo = go[t] * mo[t];
                        s[t] = fma(bo, gband[t * 16 + 8 + t], s[t]);  // leave-one-out: + beta_old |X_t|^2
                    }
#pragma unroll
                    for (int t = 0; t < kBlk; ++t) {
                        const double bo = go[t] * mo[t];
                        const double m = a * s[t];                                    // :73
                        const double x = P.c * (dd[t] - m * m * hinv + cst);          // :75-77
                        const double gm = 1.0 / (1.0 + exp(x));                       // == exp(-log1pexp(x))
                        const double bn = gm * m;                                     // :79
                        const bool live = id[t] >= 0;
                        dl[t] = live ? bn - bo : 0.0;
#pragma unroll
                        for (int u = t + 1; u < kBlk; ++u) s[u] = fma(-gband[u * 16 + 8 + t], dl[t], s[u]);
                        if (live) {
                            sg += gm;
                            sgm2 = fma(gm * m, m, sgm2);
                            sb2 = fma(bn, bn, sb2);
                            sz += fma(gm, ww[t], ii[t]);
                            if (valid) {
                                const size_t off = (size_t)id[t] * P.q_pad + k;
                                P.gam[off] = gm;
                                P.mu[off] = m;
                            }
                        }
                    }
                } else {
                    if (gb >= 2) mbar_wait(&sdone[gb & 1], (uint32_t)(((gb >> 1) - 1) & 1));  // buffer free again
#pragma unroll
                    for (int t = 0; t < kBlk; ++t) {
                        const bool live = id[t] >= 0;
                        const double bo = go[t] * mo[t];
                        dl[t] = live ? bo : 0.0;  // R = Y - X beta: subtract X_t beta_t
                        if (live) {
                            sg += go[t];
                            sgm2 = fma(go[t] * mo[t], mo[t], sgm2);
                            sb2 = fma(bo, bo, sb2);
                        }
                    }
                }
                if (active) {
                    double* db = dbuf + (size_t)(gb & 1) * kT * kBlk + tl * kBlk;
#pragma unroll
                    for (int t = 0; t < kBlk; t += 2) {
                        double2 v;
                        v.x = -dl[t];
                        v.y = -dl[t + 1];
                        *reinterpret_cast<double2*>(db + t) = v;
                    }
                }
#pragma unroll
                for (int t = 0; t < kBlk; ++t) dprev[t] = dl[t];
                __syncwarp();
                if (lane == 0) mbar_arrive(&dready[gb & 1]);
            }
            if (valid) {
                P.cs_gam[k] = sg;
                P.cs_gmu2[k] = sgm2;
                P.cs_b2[k] = sb2;
                if (P.mode == 0) P.cs_z[k] = sz;
            }
        }
    }
}

}  // namespace aq
