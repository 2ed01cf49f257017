// Shared device helpers for the atlasqtl_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace aq {

// ---------------------------------------------------------------- geometry shared by host and device
constexpr int kBlk = 8;          // SNPs per block: one m8n8k4 N-tile (S GEMM) / two K-steps (update GEMM)
constexpr int kTileTail = 144;   // doubles appended to each X tile: 128 (Gram band) + 16 (8 x int32 SNP ids, padded)

// X tile image in global memory (one per block of kBlk SNPs in sweep order), copied verbatim into
// shared memory by one bulk copy:
//   [kBlk][xs]  doubles   X columns of the block's SNPs, samples contiguous, sample index XOR-swizzled
//                         (i ^ ((t & 2) << 1)) so that both MMA operand patterns are bank-conflict free
//   [kBlk][16]  doubles   Gram band: g[t][0..7] = X_t' X_u, u in the NEXT block; g[t][8+u] = X_t' X_u, u in this block
//   [kBlk]      int32     SNP index (row of the p x q arrays) of slot t, -1 for padding slots; then 8 int32 of padding
__host__ __device__ inline size_t tile_doubles(int xs) { return (size_t)kBlk * xs + kTileTail; }
__host__ __device__ inline int swz(int i, int t) { return i ^ ((t & 2) << 1); }

// ---------------------------------------------------------------- mbarrier / bulk-copy PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Spin limit of every wait in the kernels: a protocol slip (a missed arrive, a wrong parity) traps -- the launch fails
// with an error the host reports -- instead of hanging the GPU.  A failed try_wait suspends the warp for a while, so
// 2^28 of them are seconds; no wait of a correct run is longer than a few blocks of work (microseconds).
constexpr uint32_t kSpinLimit = 1u << 28;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins == kSpinLimit) __trap();
    }
}
// 1-D bulk copy global -> shared through the TMA engine; completion is signalled on `bar` (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 16-byte asynchronous copy global -> shared (SASS: LDGSTS), tracked per thread by commit / wait groups
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------- thread-block cluster / DSMEM PTX
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_count_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* local_smem, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local_smem)), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f64(uint32_t raddr, double a) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(raddr), "d"(a) : "memory");
}
// asynchronous 16-byte store into another CTA's shared memory; its bytes are accounted (complete_tx) on that CTA's
// mbarrier `rbar`, so the receiver needs no remote arrive and the sender no fence
__device__ __forceinline__ void st_async_v2(uint32_t raddr, double a, double b, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(raddr), "d"(a),
                 "d"(b), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t raddr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {  // acquire at cluster scope
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!ok && ++spins == kSpinLimit) __trap();
    }
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- inter-CTA hand-off of a trait tile (segmented sweeps)
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---------------------------------------------------------------- fp64 tensor-core MMA (SASS: DMMA.8x8x4)
// D(8x8) += A(8x4) * B(4x8).  Lane (g = lane>>2, l = lane&3) holds A[g][l], B[l][g], D[g][2l], D[g][2l+1].
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------- chain math
// 1 / (1 + exp(x)) with a short dependency chain: the in-block Gauss-Seidel recurrence is latency bound (one fp64
// op is ~9-17 cycles behind the tensor work of the same SM sub-partition) and this routine is most of a step, so
//   - exp(r), |r| <= ln2/2, is a degree-13 Taylor polynomial in Estrin form (depth 4 instead of 13; truncation 4e-18);
//   - 2^k is assembled on the integer side while the polynomial runs, and 1 + 2^k e^r is ONE fma;
//   - the reciprocal is the hardware seed (relative error e0 <= 2^-20) times (1 + e0 + e0^2): error e0^3 < 1e-18;
//   - out-of-range arguments are resolved by selects at the end instead of clamps at the start.
// kShort trims two more dependent operations: the argument reduction r = x - k ln2 becomes ONE fma (ln2 as a double is off
// by 5.5e-17, so the result is off by the RELATIVE amount |k| 5.5e-17 -- 3e-15 for |x| <= 40, up to 6e-14 at |x| = 700
// where the value itself is 1e-304 or 1 - 1e-304: the absolute error of gam_vb stays below an ulp of 1), and the two
// range selects become one whose condition and alternative are ready long before.  Measured on B200 (same box,
// gpurun_out/r2_ab.log): chain-bound tiles gain (n = 500, 32-trait tiles: 4.35 -> 4.28 ms), the tensor-bound 16-trait tiles
// of n = 1000 lose (18.25 -> 18.45 ms), so the sweep kernel picks per configuration.
// == exp(-log1pexp(x)) of the reference (src/coreLoop.cpp:28-33, :75-77).
template <bool kShort = false>
__device__ __forceinline__ double logistic_neg(double x) {
    const double kInvLn2 = 1.4426950408889634074, kLn2 = 6.93147180559945286227e-01;
    const double kMagic = 6755399441055744.0;  // 1.5 * 2^52: round-to-nearest-integer trick
    const double t = fma(x, kInvLn2, kMagic);
    const double k = t - kMagic;
    double r;
    if constexpr (kShort) {
        r = fma(-k, kLn2, x);
    } else {
        r = fma(-k, 6.93147180369123816490e-01, x);    // ln2 in two pieces (Cody-Waite)
        r = fma(-k, 1.90821492927058770002e-10, r);
    }
    // beyond +-700 the function is 0 / 1 to 1e-304 and the pieces below are meaningless; NaN falls through (both false)
    const bool out = fabs(x) > 700.0;
    const double alt = x > 0.0 ? 0.0 : 1.0;
    const double r2 = r * r;
    const double p01 = r + 1.0;
    const double p23 = fma(r, 1.0 / 6.0, 0.5);
    const double p45 = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    const double p67 = fma(r, 1.0 / 5040.0, 1.0 / 720.0);
    const double p89 = fma(r, 1.0 / 362880.0, 1.0 / 40320.0);
    const double pab = fma(r, 1.0 / 39916800.0, 1.0 / 3628800.0);
    const double pcd = fma(r, 1.0 / 6227020800.0, 1.0 / 479001600.0);
    const double r4 = r2 * r2;
    const double q0 = fma(p23, r2, p01);
    const double q1 = fma(p67, r2, p45);
    const double q2 = fma(pab, r2, p89);
    const double r8 = r4 * r4;
    const double h0 = fma(q1, r4, q0);
    const double h1 = fma(pcd, r4, q2);
    const double e = fma(h1, r8, h0);  // exp(r)
    // 2^k: the low word of the magic sum holds the integer k (|k| <= 1010 for |x| <= 700: a normal number)
    const double scale = __hiloint2double((__double2loint(t) + 1023) << 20, 0);
    const double d = fma(e, scale, 1.0);  // 1 + exp(x) in [1, 1e304]
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e0 = fma(-d, y, 1.0);
    y = fma(y, fma(e0, e0, e0), y);
    if constexpr (kShort) return out ? alt : y;
    return x > 700.0 ? 0.0 : (x < -700.0 ? 1.0 : y);
}

}  // namespace aq
