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
    const double* xtiles;   // nb * ncta tile images, tile_stride doubles apart: image (b, r) at (b * ncta + r)
    size_t tile_stride;
    int nb;                 // number of SNP blocks = p_pad / 8
    int ntiles;             // trait tiles = q_pad / kT
    int q;                  // valid traits
    int q_pad;              // leading dimension of the p x q arrays (trait-contiguous)
    int ld_resid;           // leading dimension of resid (samples per trait row) = ncta * kNPad
    int ncta;               // CTAs per cluster = sample slices (1 for the single-CTA kernel)
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

constexpr int kMaxCluster = 8;

template <int WS_, int WT_, int MT_, int NT_, bool CL_ = false>
struct SweepCfg {
    static constexpr int WS = WS_;  // MMA warps along samples (split-K of the S GEMM)
    static constexpr int WT = WT_;  // MMA warps along traits
    static constexpr int MT = MT_;  // 8-trait M tiles per MMA warp
    static constexpr int NT = NT_;  // 8-sample N tiles per MMA warp
    static constexpr bool kCl = CL_;  // sample-split thread-block cluster variant (n > 1008)
    static constexpr int kMmaWarps = WS * WT;
    static constexpr int kT = WT * MT * 8;           // traits per tile
    static constexpr int kChainWarps = (kT + 31) / 32;
    static constexpr int kNPad = WS * NT * 8;        // samples per CTA, padded
    static constexpr int kXS = kNPad + ((kNPad % 16 == 0) ? 8 : 0);  // tile row stride == 8 (mod 16) doubles
    static constexpr int kThreads = 12 * 32;  // warps 3, 7, 11 (SMSP 3): chain warp(s) + producer
    static constexpr int kStages = 3;
    static constexpr size_t kTileDoubles = (size_t)kBlk * kXS + kTileTail;
    static constexpr int kSps = 10;  // S-partial row stride (doubles): 8 SNP slots + 2 pad => conflict-free 16-byte accesses
    static constexpr size_t kSpartDoubles = (size_t)2 * WS * kT * kSps;
    static constexpr size_t kDbufDoubles = (size_t)2 * kT * kBlk;
    static constexpr size_t kRsqDoubles = (size_t)WS * kT;
    // cluster variant only: followers' reduced S tiles and squared-norm partials land in the leader's shared memory
    static constexpr size_t kRedDoubles = kCl ? (size_t)2 * (kMaxCluster - 1) * kT * kSps : 0;
    static constexpr size_t kRsqAllDoubles = kCl ? (size_t)(kMaxCluster - 1) * kT : 0;
    static constexpr int kNumBars = 2 * kStages + 2 + 2 + 2 + 2 + 1;  // full, empty, sdone, dready, dcons, sred, rsqbar
    static constexpr size_t kSmemBytes =
        (kStages * kTileDoubles + kSpartDoubles + kDbufDoubles + kRsqDoubles + kRedDoubles + kRsqAllDoubles) * sizeof(double) +
        16 * sizeof(uint64_t);
    static_assert(kNumBars <= 16, "barrier block");
    static_assert(kMmaWarps == 9, "9 MMA warps: three per SMSP on SMSPs 0-2");
    static_assert(kChainWarps <= 2, "at most 64 traits per tile");
    static_assert(!kCl || kChainWarps == 1, "cluster variant: one chain / reducer warp");
    static_assert(kXS % 16 == 8, "row stride must be 8 mod 16 doubles");
    static_assert(kSmemBytes <= 232448, "shared memory budget (227 KB)");
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::kThreads, 1) sweep_kernel(const SweepParams P) {
    constexpr int WS = Cfg::WS, MT = Cfg::MT, NT = Cfg::NT, kT = Cfg::kT, XS = Cfg::kXS;
    constexpr int kStages = Cfg::kStages;
    constexpr bool kCl = Cfg::kCl;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);
    double* spart = tiles + kStages * Cfg::kTileDoubles;  // [2][WS][kT][kSps]
    double* dbuf = spart + Cfg::kSpartDoubles;            // [2][kT][kBlk]  (holds -Delta)
    double* rsqs = dbuf + Cfg::kDbufDoubles;              // [WS][kT]
    double* red = rsqs + Cfg::kRsqDoubles;                // leader: [2][kMaxCluster-1][kT][kSps]
    double* rsq_all = red + Cfg::kRedDoubles;             // leader: [kMaxCluster-1][kT]
    uint64_t* bars = reinterpret_cast<uint64_t*>(rsq_all + Cfg::kRsqAllDoubles);
    uint64_t* full = bars;                       // [kStages]  tile landed (tx bytes)
    uint64_t* empty = bars + kStages;            // [kStages]  MMA warps released the tile
    uint64_t* sdone = bars + 2 * kStages;        // [2]  this CTA's MMA warps wrote their S partials
    uint64_t* dready = bars + 2 * kStages + 2;   // [2]  chain warps published -Delta (in every CTA of the cluster)
    uint64_t* dcons = bars + 2 * kStages + 4;    // [2]  mode 1 only, leader: all MMA warps consumed -Delta buffer
    uint64_t* sred = bars + 2 * kStages + 6;     // [2]  leader: followers delivered their reduced S tiles
    uint64_t* rsqbar = bars + 2 * kStages + 8;   // [1]  leader: followers delivered their squared-norm partials

    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // role map: SMSP 3 (wid % 4 == 3) hosts the special warps, SMSPs 0-2 the MMA warps
    const bool is_special = (wid & 3) == 3;
    const int mma_idx = wid - (wid >> 2);   // 0..8 for MMA warps
    const int special_idx = wid >> 2;       // 0..2 for special warps
    const int ncta = kCl ? P.ncta : 1;
    const int rank = kCl ? (int)cluster_ctarank() : 0;
    const int group = kCl ? (int)cluster_id_x() : (int)blockIdx.x;         // tile-loop index of this CTA (cluster)
    const int ngroups = kCl ? (int)cluster_count_x() : (int)gridDim.x;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], Cfg::kMmaWarps); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&sdone[s], Cfg::kMmaWarps);
            mbar_init(&dready[s], Cfg::kChainWarps);
            mbar_init(&dcons[s], Cfg::kMmaWarps * ncta);
            mbar_init(&sred[s], ncta > 1 ? ncta - 1 : 1);
        }
        mbar_init(&rsqbar[0], ncta > 1 ? ncta - 1 : 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (kCl) cluster_sync_all();  // every CTA's barriers are initialised before any remote arrive

    const int nb = P.nb;
    const int my_tiles = (P.ntiles - group + ngroups - 1) / ngroups;
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
                bulk_g2s(tiles + stage * Cfg::kTileDoubles, P.xtiles + ((size_t)b * ncta + rank) * P.tile_stride, tile_bytes,
                         &full[stage]);
            }
        }
    } else if (!is_special) {
        // ------------------------------------------------------------------ MMA warps
        const int ws = mma_idx % WS, wt = mma_idx / WS;
        const int mtid = mma_idx * 32 + lane;
        const int g = lane >> 2, l = lane & 3;
        const int i0 = ws * NT * 8;                       // first sample of this warp inside the CTA's slice
        const int ig = rank * Cfg::kNPad + i0;            // ... and inside the residual row
        const int tr0 = wt * MT * 8;
        // lane-constant shared-memory offsets of the two operand patterns
        const int offS = g * XS + ((i0 + 2 * l) ^ ((g & 2) << 1));         // + nt*8   (16-byte loads)
        const int offU0 = l * XS + ((i0 + g) ^ ((l & 2) << 1));            // ks = 0, + nt*8
        const int offU1 = (l + 4) * XS + ((i0 + g) ^ ((l & 2) << 1));      // ks = 1 (snp l+4 has the same bit 1)
        uint32_t dcons_leader[2] = {0, 0}, rsqbar_leader = 0, rsq_all_leader = 0;
        if (kCl) {
            dcons_leader[0] = mapa_u32(&dcons[0], 0);
            dcons_leader[1] = mapa_u32(&dcons[1], 0);
            rsqbar_leader = mapa_u32(&rsqbar[0], 0);
            rsq_all_leader = mapa_u32(rsq_all, 0);
        }
        double acc[MT][NT][2];
        long gb = 0;  // global block counter of this CTA (drives ring stage and barrier parity)
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int tile = group + ti * ngroups;
            const int k0 = tile * kT;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double2 v = *reinterpret_cast<const double2*>(
                        P.resid + (size_t)(k0 + tr0 + mt * 8 + g) * P.ld_resid + ig + nt * 8 + 2 * l);
                    acc[mt][nt][0] = v.x;
                    acc[mt][nt][1] = v.y;
                }
            auto s_phase = [&](long gbi) {
                const int stage = (int)(gbi % kStages);
                mbar_wait(&full[stage], (uint32_t)((gbi / kStages) & 1));
                const double* xt = tiles + stage * Cfg::kTileDoubles;
                // kSC accumulator chains per M tile: a dependent DMMA issues every ~26 cycles, the pipe takes one per 16
                constexpr int kSC = (MT >= 4) ? 1 : 2;
                double sa[MT][2][2];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) sa[mt][0][0] = sa[mt][0][1] = sa[mt][1][0] = sa[mt][1][1] = 0.0;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double2 xb = *reinterpret_cast<const double2*>(xt + offS + nt * 8);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        dmma(sa[mt][0][0], sa[mt][0][1], acc[mt][nt][0], xb.x);
                        dmma(sa[mt][kSC - 1][0], sa[mt][kSC - 1][1], acc[mt][nt][1], xb.y);
                    }
                }
                double* sp = spart + ((size_t)(gbi & 1) * WS + ws) * kT * Cfg::kSps;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    double2 v;  // C fragment: S^T[trait g + 8 mt][snp 2l, 2l + 1]
                    v.x = (kSC == 2) ? sa[mt][0][0] + sa[mt][1][0] : sa[mt][0][0];
                    v.y = (kSC == 2) ? sa[mt][0][1] + sa[mt][1][1] : sa[mt][0][1];
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
                if (kCl) mbar_wait_cluster(&dready[gb & 1], (uint32_t)((gb >> 1) & 1));
                else mbar_wait(&dready[gb & 1], (uint32_t)((gb >> 1) & 1));
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
                    if (lane == 0) {
                        if (kCl) mbar_arrive_cluster(dcons_leader[gb & 1]);
                        else mbar_arrive(&dcons[gb & 1]);
                    }
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
                    *reinterpret_cast<double2*>(P.resid + (size_t)(k0 + tr0 + mt * 8 + g) * P.ld_resid + ig + nt * 8 + 2 * l) = v;
                }
                ss += __shfl_xor_sync(0xffffffffu, ss, 1);
                ss += __shfl_xor_sync(0xffffffffu, ss, 2);
                if (l == 0) rsqs[ws * kT + tr0 + mt * 8 + g] = ss;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(Cfg::kMmaWarps * 32) : "memory");
            if (mtid < 32 * ((kT + 31) / 32)) {  // whole warps, so that the elected arrive below is warp-uniform
                double ss = 0.0;
                if (mtid < kT) {
#pragma unroll
                    for (int w2 = 0; w2 < WS; ++w2) ss += rsqs[w2 * kT + mtid];
                }
                if (!kCl) {
                    if (mtid < kT && k0 + mtid < P.q) P.rsq[k0 + mtid] = ss;
                } else if (rank != 0) {
                    if (mtid < kT) st_cluster_f64(rsq_all_leader + (uint32_t)(((rank - 1) * kT + mtid) * sizeof(double)), ss);
                    fence_cluster();
                    __syncwarp();
                    if (mtid == 0) mbar_arrive_cluster(rsqbar_leader);
                } else {
                    if (ncta > 1) mbar_wait_cluster(&rsqbar[0], (uint32_t)(ti & 1));
                    if (mtid < kT) {
                        for (int r2 = 1; r2 < ncta; ++r2) ss += rsq_all[(r2 - 1) * kT + mtid];  // fixed order: deterministic
                        if (k0 + mtid < P.q) P.rsq[k0 + mtid] = ss;
                    }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(Cfg::kMmaWarps * 32) : "memory");
        }
    } else if (special_idx < Cfg::kChainWarps && rank != 0) {
        // ------------------------------------------------------------------ follower CTA: S-tile reducer warp
        // Sums this CTA's WS split-K partials and ships the kT x 8 tile into the leader's shared memory.
        if (P.mode == 0) {
            const int tl = special_idx * 32 + lane;
            const bool active = tl < kT;
            const int tls = active ? tl : 0;
            const uint32_t red_leader = mapa_u32(red, 0);
            const uint32_t sred_leader[2] = {mapa_u32(&sred[0], 0), mapa_u32(&sred[1], 0)};
            const long total = (long)my_tiles * nb;
            for (long gb = 0; gb < total; ++gb) {
                mbar_wait(&sdone[gb & 1], (uint32_t)((gb >> 1) & 1));
                const double* sp = spart + (size_t)(gb & 1) * WS * kT * Cfg::kSps + tls * Cfg::kSps;
                double s[kBlk];
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
                if (active) {
                    const uint32_t dst = red_leader + (uint32_t)(((((gb & 1) * (kMaxCluster - 1)) + (rank - 1)) * kT + tl) *
                                                                 Cfg::kSps * sizeof(double));
#pragma unroll
                    for (int t = 0; t < kBlk; t += 2) st_cluster_v2(dst + t * (uint32_t)sizeof(double), s[t], s[t + 1]);
                }
                fence_cluster();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(sred_leader[gb & 1]);
            }
        }
    } else if (special_idx < Cfg::kChainWarps) {
        // ------------------------------------------------------------------ chain warps (one lane per trait; leader CTA)
        const int cw = special_idx;
        const int tl = cw * 32 + lane;
        const bool active = tl < kT;
        const int tls = active ? tl : 0;  // inactive lanes shadow trait 0 without side effects
        uint32_t dbuf_remote[kMaxCluster], dready_remote[kMaxCluster][2];
        if (kCl) {
            for (int r2 = 1; r2 < ncta; ++r2) {
                dbuf_remote[r2] = mapa_u32(dbuf, r2);
                dready_remote[r2][0] = mapa_u32(&dready[0], r2);
                dready_remote[r2][1] = mapa_u32(&dready[1], r2);
            }
        }
        long gb = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int tile = group + ti * ngroups;
            const int k = tile * kT + tls;
            const bool valid = active && k < P.q;
            const double sig2 = P.sig2_beta[k], tauk = P.tau[k];
            const double a = P.c * sig2 * tauk;                                   // src/coreLoop.cpp:73
            const double hinv = 1.0 / (2.0 * sig2);                               // :76
            const double cst = -(P.log_tau[k] + P.log_sig2_inv + log(sig2)) / 2;  // :56
            double sg = 0.0, sgm2 = 0.0, sb2 = 0.0, sz = 0.0;
            for (int b = 0; b < nb; ++b, ++gb) {
                const int stage = (int)(gb % kStages);
                mbar_wait(&full[stage], (uint32_t)((gb / kStages) & 1));
                const double* xt = tiles + stage * Cfg::kTileDoubles;
                const double* gband = xt + kBlk * XS;
                const int* ids = reinterpret_cast<const int*>(gband + 128);
                // all 40 loads are issued back to back (no use in between), so one DRAM/L2 round trip covers the block
                double go[kBlk], mo[kBlk], dd[kBlk], ww[kBlk], ii[kBlk];
#pragma unroll
                for (int t = 0; t < kBlk; ++t) {
                    const int idt = ids[t];
                    const size_t off = (size_t)(idt < 0 ? 0 : idt) * P.q_pad + k;
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
                // the previous block's -Delta row is read back from shared memory for the look-ahead correction
                double* drow = dbuf + (size_t)(gb & 1) * kT * kBlk + tls * kBlk;
                const double* dprow = dbuf + (size_t)((gb + 1) & 1) * kT * kBlk + tls * kBlk;
                double s[kBlk], nd[kBlk];
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
                    if (kCl && ncta > 1) {  // + the other sample slices, already reduced by their CTAs
                        mbar_wait_cluster(&sred[gb & 1], (uint32_t)((gb >> 1) & 1));
                        for (int r2 = 1; r2 < ncta; ++r2) {
                            const double* rp = red + ((size_t)((gb & 1) * (kMaxCluster - 1) + (r2 - 1)) * kT + tls) * Cfg::kSps;
#pragma unroll
                            for (int t = 0; t < kBlk; t += 2) {
                                const double2 v = *reinterpret_cast<const double2*>(rp + t);
                                s[t] += v.x;
                                s[t + 1] += v.y;
                            }
                        }
                    }
                    // look-ahead correction: S was formed before the previous block's update was applied
                    if (b > 0) {
#pragma unroll
                        for (int u = 0; u < kBlk; u += 2) {
                            const double2 nd2 = *reinterpret_cast<const double2*>(dprow + u);  // -Delta_prev[u], [u+1]
#pragma unroll
                            for (int t = 0; t < kBlk; ++t) {
                                s[t] = fma(gband[t * 16 + u], nd2.x, s[t]);
                                s[t] = fma(gband[t * 16 + u + 1], nd2.y, s[t]);
                            }
                        }
                    }
                    double sum_i0 = 0.0;
#pragma unroll
                    for (int t = 0; t < kBlk; ++t) {
                        const bool live = ids[t] >= 0;
                        go[t] = live ? go[t] * mo[t] : 0.0;  // beta_old (0 for padding slots: their X column is 0)
                        s[t] = fma(go[t], gband[t * 16 + 8 + t], s[t]);  // leave-one-out: + beta_old |X_t|^2
                        sum_i0 += live ? ii[t] : 0.0;
                    }
#pragma unroll
                    for (int t = 0; t < kBlk; ++t) {
                        const double m = a * s[t];                                    // :73
                        const double x = P.c * (dd[t] - m * m * hinv + cst);          // :75-77
                        const double gm = logistic_neg(x);                            // 1/(1+e^x) == exp(-log1pexp(x))
                        const double bn = gm * m;                                     // :79
                        const double dlt = bn - go[t];                                // 0 for padding slots (m == 0)
#pragma unroll
                        for (int u = t + 1; u < kBlk; ++u) s[u] = fma(-gband[u * 16 + 8 + t], dlt, s[u]);
                        nd[t] = -dlt;
                        const int idt = ids[t];
                        if (idt >= 0) {
                            sg += gm;
                            sgm2 = fma(gm * m, m, sgm2);
                            sb2 = fma(bn, bn, sb2);
                            sz = fma(gm, ww[t], sz);
                            if (valid) {
                                const size_t off = (size_t)idt * P.q_pad + k;
                                P.gam[off] = gm;
                                P.mu[off] = m;
                            }
                        }
                    }
                    sz += sum_i0;
                } else {
                    if (gb >= 2) {  // -Delta buffer (gb & 1) free again in every CTA
                        if (kCl) mbar_wait_cluster(&dcons[gb & 1], (uint32_t)(((gb >> 1) - 1) & 1));
                        else mbar_wait(&dcons[gb & 1], (uint32_t)(((gb >> 1) - 1) & 1));
                    }
#pragma unroll
                    for (int t = 0; t < kBlk; ++t) {
                        const bool live = ids[t] >= 0;
                        const double bo = live ? go[t] * mo[t] : 0.0;  // R = Y - X beta: subtract X_t beta_t
                        if (live) {
                            sg += go[t];
                            sgm2 = fma(go[t] * mo[t], mo[t], sgm2);
                            sb2 = fma(bo, bo, sb2);
                        }
                        nd[t] = -bo;
                    }
                }
                // publish -Delta only now: a shared-memory store inside the step loop would order every later Gram-band
                // load behind it (possible aliasing) and put the LDS latency on the serial path of each step
                if (active) {
#pragma unroll
                    for (int t = 0; t < kBlk; t += 2) {
                        double2 v;
                        v.x = nd[t];
                        v.y = nd[t + 1];
                        *reinterpret_cast<double2*>(drow + t) = v;
                        if (kCl)
                            for (int r2 = 1; r2 < ncta; ++r2)
                                st_cluster_v2(dbuf_remote[r2] + (uint32_t)(((size_t)(gb & 1) * kT * kBlk + tl * kBlk + t) * sizeof(double)), v.x, v.y);
                    }
                }
                if (kCl) fence_cluster();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&dready[gb & 1]);
                    if (kCl)
                        for (int r2 = 1; r2 < ncta; ++r2) mbar_arrive_cluster(dready_remote[r2][gb & 1]);
                }
            }
            if (valid) {
                P.cs_gam[k] = sg;
                P.cs_gmu2[k] = sgm2;
                P.cs_b2[k] = sb2;
                if (P.mode == 0) P.cs_z[k] = sz;
            }
        }
    }
    if (kCl) {
        __syncwarp();
        cluster_sync_all();  // keep every CTA's shared memory alive until all remote stores / arrives have landed
    }
}

}  // namespace aq
