// tcgen05 / TMEM / mbarrier helpers and the round-1 3xTF32 configuration shared by tc_persist.cuh (ADEC_CONV_PATH=tf32, kept for A/B)
// and tc_f16.cuh (default).  The one-tile-per-CTA tc_conv_kernel of round 1 was removed in round 2 (the persistent schedule
// superseded it: 18.95 -> 14.03 ms per step, profiles/README.md).  Design notes of the 3xTF32 engine:
//
// Same ConvArgs contract as conv_gemm_kernel (kernels.cuh); different engine:
//   * GEMM orientation: M = 128 time steps (TMEM lanes), N = output channels (TMEM columns), K = input
//     channels of one tap.  A (activations) and B (weights) are both K-major, no-swizzle UMMA operands
//     stored as "column blocks"  smem[(k/4)*ROWS + row][k%4]  (16-byte core-matrix rows, SBO = 128 B,
//     LBO = ROWS*16 B).  In this layout a conv tap is a *row-shifted start address* of the same
//     window - no im2col, no copies: tap k of a dilated conv reads rows [k*dil, k*dil+128).
//   * precision: 3xTF32 error-compensated products (a = a_hi + a_lo, w = w_hi + w_lo in tf32;
//     a_lo*w_hi + a_hi*w_lo + a_hi*w_hi).  The TMEM accumulator rounds toward zero at every MMA
//     (measured, tools/tc_probe2.cu: -1.7e-8..-3.8e-8 relative per accumulation step), which over the
//     hundreds of steps of a long-K conv becomes a 1e-5-level systematic shrink - enough to flip
//     nearest-codeword decisions.  So accumulation is GROUPED: one (32-channel piece, tap) = 12 MMAs
//     (8 small-term MMAs first, then 4 main ones) goes into a fresh TMEM partial, and the drain warps
//     add the partials into fp32 REGISTER accumulators with round-to-nearest adds.  Residual bias
//     ~1e-7 per conv, same order as the FFMA path's rounding noise.
//   * warp roles (384 or 512 threads): warp 0 = weight producer (cp.async.bulk / TMA 1-D; host-pre-split
//     hi|lo tiles already in UMMA layout) + TMEM allocator; warp 1 = MMA issuer (elect-one, uniform
//     datapath); warps 4-7 = activation producers (global -> pre-activation -> hi/lo split -> smem,
//     32-channel pieces, double buffered); warps 8+ = drain / epilogue (tcgen05.ld partials, register
//     accumulation, bias / residual, stores).  mbarrier rings connect them; tcgen05.commit releases smem
//     stages and signals partials.
//   * FUSE (C <= 128): the residual unit's activated intermediate goes from the drain warps' registers
//     straight back to smem as the hi/lo operand of the 1x1 conv; the skip tensor is added in the epilogue.
#pragma once
#include "kernels.cuh"

namespace adec {

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes) {
    // K-major, SWIZZLE_NONE: start>>4 | LBO>>4 <<16 | SBO(=128 B)>>4 <<32 | version 1 <<46
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) |
           ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// commit that arrives on the same barrier offset in every CTA of `mask` (weights are shared by the cluster)
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
// bulk copy whose bytes (and mbarrier complete_tx) land at the same CTA-relative offsets in every CTA of `mask`
__device__ __forceinline__ void bulk_g2s_mc(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n .reg .pred P1;\n elect.sync _|P1, 0xffffffff;\n selp.u32 %0, 1, 0, P1;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// 32 lanes x 16 columns of fp32 from TMEM (lane quarter of this warp), no wait
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

constexpr int TC_TT = 128;        // output rows per CTA (UMMA M)
constexpr int TC_CP = 32;         // channels per activation piece == K of one weight stage
constexpr int TC_MIDP = 129;      // row pitch (rows) of the 1x1 conv's activated operand (odd)

template <int NT>
struct TcCfg {
    static constexpr int STAGES = NT == 64 ? 4 : 3;
    static constexpr int MIN_CTAS = NT == 32 ? 2 : 1;
    static constexpr int CLUSTER = 1;                       // CTAs (adjacent time tiles) sharing one weight stream                       // co-resident CTAs hide each other's serial phases
    static constexpr int B_STAGE_FLOATS = 2 * TC_CP * NT;                   // hi | lo
    static constexpr int NDG = NT >= 64 ? 2 : 1;                            // drain warpgroups (each owns NT/NDG columns)
    static constexpr int NPROD = NT == 128 ? 128 : 256;                     // activation-producer threads
    static constexpr int DRAIN0 = (128 + NPROD) / 32;                       // first drain warp
    static constexpr int THREADS = 128 + NPROD + 128 * NDG;                 // warps 0-3 control, then producers, then drain
    static constexpr int NCOL = NT / NDG;                                   // accumulator registers per drain thread
};

template <int ACT>
__device__ __forceinline__ float4 apply_act_t(float4 v, float slope) {
    if (ACT == ACT_ELU) { v.x = act_elu(v.x); v.y = act_elu(v.y); v.z = act_elu(v.z); v.w = act_elu(v.w); }
    if (ACT == ACT_LRELU) { v.x = act_lrelu(v.x, slope); v.y = act_lrelu(v.y, slope); v.z = act_lrelu(v.z, slope); v.w = act_lrelu(v.w, slope); }
    return v;
}

}  // namespace adec
