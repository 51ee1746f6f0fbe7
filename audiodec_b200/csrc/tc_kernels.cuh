// tc_conv_kernel: the causal conv family on 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
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

// Two MMA warps take alternate groups c = 0,1,2,... of one weight ring with S stages.  When S is odd, consecutive phases of a
// stage's b_full barrier belong to DIFFERENT warps, and an mbarrier parity wait only tells "phase k complete" from "phase k
// pending" for a waiter that knows phase k-1 has completed (otherwise a pending phase k-1 reads as a completed phase k: the warp
// would issue MMAs on a stage whose TMA load has not landed and release it early; seen as rare hangs when two weight loads
// landed out of order).  So each warp publishes the last group whose weights it has seen land, and before waiting for group c a
// warp makes sure the other one has seen group c - S, the previous phase of the same stage.  With S even a stage always belongs
// to the same warp and nothing is needed.
template <int S> __device__ __forceinline__ void mma_wait_turn(volatile int* prog, int mw, int c) {
    if ((S & 1) && c >= S) {
        long long t0 = 0;
        unsigned spins = 0;
        while (prog[mw ^ 1] < c - S) {
            __nanosleep(64);        // a hot spin here steals issue slots from the producer/drain warps of the same sub-partition (measured: -10..20 %)
            if ((++spins & 0xfffu) == 0) {
                if (t0 == 0) t0 = clock64();
                else if (clock64() - t0 > 4000000000ll) { printf("adec: MMA warp order wait timed out: group %d\n", c); __trap(); }
            }
        }
    }
}
template <int S> __device__ __forceinline__ void mma_publish(volatile int* prog, int mw, int c, int lane) {
    if (S & 1) { if (lane == 0) prog[mw] = c; }
}

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

// PRE: pre-activation applied to chunk rows while the window is written (also the residual unit's mid activation)
template <int NT, bool FUSE, int PRE>
__global__ void __launch_bounds__(TcCfg<NT>::THREADS, TcCfg<NT>::MIN_CTAS) tc_conv_kernel(const ConvArgs a) {
    using Cfg = TcCfg<NT>;
    constexpr int S = Cfg::STAGES, BST = Cfg::B_STAGE_FLOATS, CP = TC_CP, TT = TC_TT, NDG = Cfg::NDG, NCOL = Cfg::NCOL;
    constexpr int NPROD = Cfg::NPROD, DRAIN0 = Cfg::DRAIN0;
    constexpr int MIDP = TC_MIDP;
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t TMEM_COLS = 2 * NT;      // two partial-accumulator buffers

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* b_full = reinterpret_cast<uint64_t*>(smem_raw);     // [S]  weights landed
    uint64_t* b_empty = b_full + S;                                // [S]  weights consumed
    uint64_t* a_full = b_empty + S;                                // [2]  activation piece written
    uint64_t* a_empty = a_full + 2;                                // [2]  activation piece consumed
    uint64_t* p_full = a_empty + 2;                                // [2]  TMEM partial complete
    uint64_t* p_empty = p_full + 2;                                // [2]  TMEM partial drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_empty + 2);
    volatile int* mma_prog = reinterpret_cast<volatile int*>(smem_raw + 240);   // [2] see mma_wait_turn
    float* bst = reinterpret_cast<float*>(smem_raw + 256);
    const int wrows = TT + (a.Ktaps - 1) * a.dil;
    const int wrp = (wrows > MIDP ? wrows : MIDP) | 1;             // odd row pitch: conflict-free producer stores
    float* abuf0 = bst + S * BST;
    float* abuf1 = abuf0 + 2 * CP * wrp;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);       // warp-uniform for the compiler
    const int j0 = blockIdx.x * TT;
    const int g = blockIdx.y / a.n_co_tiles;
    const int co_tile = blockIdx.y - g * a.n_co_tiles;
    const int b = blockIdx.z;
    const int n_g1 = a.n_pieces * a.Ktaps;                          // groups (= weight stages) of GEMM 1
    const int n_g2 = FUSE ? NT / CP : 0;                            // groups of the fused 1x1 conv

    // Clusters along the time axis share the weight stream: every CTA fetches 1/CL of each stage and multicasts it.
    const uint32_t CL = cluster_nctarank(), crank = cluster_ctarank();
    const uint16_t cmask = (uint16_t)((1u << CL) - 1u);
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], CL); }   // freed when ALL CTAs consumed it
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], NPROD); mbar_init(&a_empty[i], 2);   // both MMA warps release a piece
            mbar_init(&p_full[i], 1); mbar_init(&p_empty[i], 128 * NDG);
        }
        mma_prog[0] = -1; mma_prog[1] = -1;
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();      // peers' barriers are initialised before anyone multicasts into them
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
#ifdef ADEC_TIMELINE
    __shared__ unsigned tl_[6][64];
    const long long tl0_ = clock64();
#define TL(role, i) do { if ((i) < 64) tl_[role][(i)] = (unsigned)(clock64() - tl0_); } while (0)
#else
#define TL(role, i) do { } while (0)
#endif

    if (warp == 0) {
        // ------------------------------------------------ weight producer (TMA 1-D bulk copies)
        if (lane == 0) {
            const float* w1 = a.w + (long long)blockIdx.y * a.w_tile_floats;
            for (int c = 0; c < n_g1 + n_g2; ++c) {
                const int s = c % S, it = c / S;
                if (it > 0) mbar_wait(&b_empty[s], (it - 1) & 1, 100 + c);
                const float* src = c < n_g1 ? w1 + (long long)c * BST : a.w2 + (long long)(c - n_g1) * BST;
                mbar_arrive_expect_tx(&b_full[s], BST * 4);
                if (CL == 1) {
                    bulk_g2s(bst + s * BST, src, BST * 4, &b_full[s]);
                } else {
                    const uint32_t slice = (uint32_t)BST / CL;      // floats
                    bulk_g2s_mc(bst + s * BST + crank * slice, src + crank * slice, slice * 4, &b_full[s], cmask);
                }
                TL(0, c);
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ------------------------------------------------ MMA issuers: two warps take alternate groups (= alternate TMEM
        // partials), so one is already past its barrier waits when the other's MMAs leave the queue.  Whole warp runs the
        // loop (uniform datapath), one elected lane issues.
        const int mw = warp - 1;
        int c = 0, gp = 0;
        const uint32_t b_lbo = (uint32_t)NT * 16u;
        const uint32_t abuf0_u = smem_u32(abuf0), abuf1_u = smem_u32(abuf1), bst_u = smem_u32(bst);
        for (int phase = 0; phase < (FUSE ? 2 : 1); ++phase) {
            const int pieces = phase == 0 ? a.n_pieces : NT / CP;
            const int taps = phase == 0 ? a.Ktaps : 1;
            const uint32_t lbo = (uint32_t)(phase == 0 ? wrp : MIDP) * 16u;
            for (int p = 0; p < pieces; ++p, ++gp) {
                const int buf = gp & 1;
                mbar_wait(&a_full[buf], (gp >> 1) & 1, 200 + gp);
                const uint32_t a_hi = buf ? abuf1_u : abuf0_u;
                const uint32_t a_lo = a_hi + (uint32_t)(CP / 4) * lbo;     // lo block follows the hi block
                for (int tap = 0; tap < taps; ++tap, ++c) {
                    if ((c & 1) != mw) continue;
                    const int s = c % S, pb = c & 1;
                    mma_wait_turn<S>(mma_prog, mw, c);
                    mbar_wait(&b_full[s], (c / S) & 1, 300 + c);
                    mma_publish<S>(mma_prog, mw, c, lane);
                    if (c >= 2) mbar_wait(&p_empty[pb], ((c >> 1) - 1) & 1, 400 + c);
                    tc_fence_after();
                    if (lane == 0) TL(1, c);
                    const uint32_t row_off = (uint32_t)(tap * a.dil) * 16u;
                    const uint32_t b_hi = bst_u + (uint32_t)s * (BST * 4u);
                    const uint32_t b_lo = b_hi + (uint32_t)CP * NT * 4u;
                    const uint32_t acc = tmem + (uint32_t)pb * NT;
                    if (elect_one()) {
                        // small terms first (the partial is still tiny while they accumulate), then the main products
#pragma unroll
                        for (int k8 = 0; k8 < CP / 8; ++k8)
                            umma_tf32(acc, umma_desc(a_lo + (uint32_t)(k8 * 2) * lbo + row_off, lbo),
                                      umma_desc(b_hi + (uint32_t)(k8 * 2) * b_lbo, b_lbo), IDESC, k8 ? 1u : 0u);
#pragma unroll
                        for (int k8 = 0; k8 < CP / 8; ++k8)
                            umma_tf32(acc, umma_desc(a_hi + (uint32_t)(k8 * 2) * lbo + row_off, lbo),
                                      umma_desc(b_lo + (uint32_t)(k8 * 2) * b_lbo, b_lbo), IDESC, 1u);
#pragma unroll
                        for (int k8 = 0; k8 < CP / 8; ++k8)
                            umma_tf32(acc, umma_desc(a_hi + (uint32_t)(k8 * 2) * lbo + row_off, lbo),
                                      umma_desc(b_hi + (uint32_t)(k8 * 2) * b_lbo, b_lbo), IDESC, 1u);
                        if (CL == 1) umma_commit(&b_empty[s]);       // weight stage free once these MMAs retire
                        else umma_commit_mc(&b_empty[s], cmask);     // ... in every CTA of the cluster
                        umma_commit(&p_full[pb]);       // partial ready for the drain warps
                        TL(2, c);
                    }
                    __syncwarp();
                }
                if (elect_one()) umma_commit(&a_empty[buf]);   // this warp's MMAs on the piece (if any) have retired
                __syncwarp();
            }
        }
    } else if (warp >= 4 && warp < DRAIN0) {
        // ------------------------------------------------ activation producers (NPROD threads)
        const int pt = tid - 128;
        const float* xg = a.x + (long long)b * a.x_bs + g * a.x_goff;
        const float* sg = a.st_in + (long long)b * a.P * a.st_ld + g * a.st_goff;
        for (int p = 0; p < a.n_pieces; ++p) {
            const int buf = p & 1;
            if (p >= 2) mbar_wait(&a_empty[buf], ((p >> 1) - 1) & 1, 500 + p);
            float* hi = buf ? abuf1 : abuf0;
            float* lo = hi + CP * wrp;
            // thread -> fixed 4-channel column c4 and rows m0, m0+RPP, ...: all index math is loop-invariant
            constexpr int RPP = NPROD / 8;               // rows per pass
            constexpr int UNR = 6;                       // loads in flight per thread (memory-level parallelism)
            const int c4 = pt & 7, m0 = pt >> 3;
            const int q = p * CP + c4 * 4;
            int r = 0, ci = q;
            if (a.RG > 1) { r = q >> a.lgCin; ci = q & (a.Cin - 1); }
            const float* srow = sg + ci;
            const float* xrow = xg + ci;
            float* hcol = hi + (c4 * wrp) * 4;
            float* lcol = lo + (c4 * wrp) * 4;
            // interior tiles (no history rows, no rows past the chunk end) take the unchecked path
            const long long i_first = (long long)j0 * a.RG + r;
            const long long i_last = (long long)(j0 + wrows - 1) * a.RG + r;
            if (i_first >= a.P && i_last - a.P < a.T && PRE != ACT_NORM) {
                const float* xp = xrow + (i_first - a.P + (long long)m0 * a.RG) * a.ldx;
                const long long xstep = (long long)RPP * a.RG * a.ldx;
                for (int mb = m0; mb < wrows; mb += RPP * UNR, xp += xstep * UNR) {
                    float4 v[UNR];
#pragma unroll
                    for (int u = 0; u < UNR; ++u)
                        if (mb + u * RPP < wrows) v[u] = __ldg(reinterpret_cast<const float4*>(xp + u * xstep));
#pragma unroll
                    for (int u = 0; u < UNR; ++u) {
                        const int m = mb + u * RPP;
                        if (m < wrows) {
                            const float4 x4 = apply_act_t<PRE>(v[u], a.slope);
                            const float4 h = make_float4(tf32_rna(x4.x), tf32_rna(x4.y), tf32_rna(x4.z), tf32_rna(x4.w));
                            const float4 l = make_float4(x4.x - h.x, x4.y - h.y, x4.z - h.z, x4.w - h.w);
                            *reinterpret_cast<float4*>(hcol + m * 4) = h;
                            *reinterpret_cast<float4*>(lcol + m * 4) = l;
                        }
                    }
                }
            } else {
                for (int mb = m0; mb < wrows; mb += RPP * UNR) {
                    float4 v[UNR];
    #pragma unroll
                    for (int u = 0; u < UNR; ++u) {
                        const int m = mb + u * RPP;
                        const long long i = (long long)(j0 + m) * a.RG + r;
                        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (m < wrows) {
                            long long ti = i - a.P;
                            if (a.hist_rep && ti < 0) ti = 0;              // non-streaming transposed conv: replicate the first input row
                            if (ti < 0) v[u] = *reinterpret_cast<const float4*>(srow + i * a.st_ld);
                            else if (ti < a.T) v[u] = __ldg(reinterpret_cast<const float4*>(xrow + ti * a.ldx));
                        }
                    }
    #pragma unroll
                    for (int u = 0; u < UNR; ++u) {
                        const int m = mb + u * RPP;
                        if (m < wrows) {
                            const long long i = (long long)(j0 + m) * a.RG + r;
                            float4 x4 = v[u];
                            if ((i >= a.P || a.hist_rep) && i - a.P < a.T) {        // chunk rows get the pre-activation; history rows already have it
                                if (PRE == ACT_NORM) {
                                    const float4 mu = *reinterpret_cast<const float4*>(a.mean + ci);
                                    const float4 sc = *reinterpret_cast<const float4*>(a.scale + ci);
                                    x4.x = __fdiv_rn(x4.x - mu.x, sc.x); x4.y = __fdiv_rn(x4.y - mu.y, sc.y);
                                    x4.z = __fdiv_rn(x4.z - mu.z, sc.z); x4.w = __fdiv_rn(x4.w - mu.w, sc.w);
                                } else {
                                    x4 = apply_act_t<PRE>(x4, a.slope);
                                }
                            }
                            const float4 h = make_float4(tf32_rna(x4.x), tf32_rna(x4.y), tf32_rna(x4.z), tf32_rna(x4.w));
                            const float4 l = make_float4(x4.x - h.x, x4.y - h.y, x4.z - h.z, x4.w - h.w);   // exact; the MMA reads its top 19 bits
                            *reinterpret_cast<float4*>(hcol + m * 4) = h;
                            *reinterpret_cast<float4*>(lcol + m * 4) = l;
                        }
                    }
                }
            }
            fence_async_smem();
            mbar_arrive(&a_full[buf]);
            if (pt == 0) TL(3, p);
        }
        // ---- new causal state (conv_layer.py:155), independent of the MMA pipeline
        if ((int)blockIdx.x == (a.Tout - 1) / TT && co_tile == 0 && g < a.st_groups && a.P > 0) {   // grid.x may be padded to the cluster size
            float* so = a.st_out + (long long)b * a.P * a.st_ld + g * a.st_goff;
            const int nvec = a.P * (a.Cin / 4);
            for (int idx = pt; idx < nvec; idx += NPROD) {
                const int r = idx / (a.Cin / 4);
                const int ci = (idx - r * (a.Cin / 4)) * 4;
                const long long i = (long long)a.T + r;
                float4 v;
                if (i < a.P) {
                    v = *reinterpret_cast<const float4*>(sg + i * a.st_ld + ci);
                } else {
                    v = __ldg(reinterpret_cast<const float4*>(xg + (i - a.P) * a.ldx + ci));
                    if (PRE == ACT_NORM) {
                        const float4 mu = *reinterpret_cast<const float4*>(a.mean + ci);
                        const float4 sc = *reinterpret_cast<const float4*>(a.scale + ci);
                        v.x = __fdiv_rn(v.x - mu.x, sc.x); v.y = __fdiv_rn(v.y - mu.y, sc.y);
                        v.z = __fdiv_rn(v.z - mu.z, sc.z); v.w = __fdiv_rn(v.w - mu.w, sc.w);
                    } else {
                        v = apply_act_t<PRE>(v, a.slope);
                    }
                }
                *reinterpret_cast<float4*>(so + (long long)r * a.st_ld + ci) = v;
            }
        }
    } else if (warp >= DRAIN0) {
        // ------------------------------------------------ drain warps: register accumulation, mid conversion, epilogue
        const int dg = (warp - DRAIN0) >> 2;                     // drain group: owns columns [dg*NCOL, (dg+1)*NCOL)
        const int row = (warp & 3) * 32 + lane;             // TMEM lane == output row of this thread
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        float racc[NCOL];
#pragma unroll
        for (int i = 0; i < NCOL; ++i) racc[i] = 0.f;
        int c = 0;
        auto drain = [&](int ngroups) {
            for (int gi = 0; gi < ngroups; ++gi, ++c) {
                const int pb = c & 1;
                mbar_wait(&p_full[pb], (c >> 1) & 1, 600 + c);
                tc_fence_after();
                if (tid == DRAIN0 * 32) TL(4, c);
                const uint32_t taddr = tmem + lane_base + (uint32_t)pb * NT + (uint32_t)dg * NCOL;
#pragma unroll
                for (int c0 = 0; c0 < NCOL; c0 += 32) {
                    uint32_t r0[16], r1[16];
                    tmem_ld16(taddr + c0, r0);
                    tmem_ld16(taddr + c0 + 16, r1);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        racc[c0 + i] = __fadd_rn(racc[c0 + i], __uint_as_float(r0[i]));
                        racc[c0 + 16 + i] = __fadd_rn(racc[c0 + 16 + i], __uint_as_float(r1[i]));
                    }
                }
                tc_fence_before();
                mbar_arrive(&p_empty[pb]);
                if (tid == DRAIN0 * 32) TL(5, c);
            }
        };
        drain(n_g1);
        if (FUSE) {
            // mid = act(conv_k7(act(x))) -> hi/lo operand of the 1x1 conv, written straight from registers
            // Piece gq re-uses the buffer of piece gq-2, i.e. needs completion #((gq>>1)-1) of a_empty[buf].  A parity
            // wait is unambiguous here because completion #((gq>>1)-2) (piece gq-4) retired long before GEMM 1's last
            // partial, which these threads have already drained.
#pragma unroll
            for (int pl = 0; pl < NCOL / CP; ++pl) {
                const int p = dg * (NCOL / CP) + pl;        // piece index in MMA consumption order
                const int gq = a.n_pieces + p, buf = gq & 1;
                if (gq >= 2) mbar_wait(&a_empty[buf], ((gq >> 1) - 1) & 1, 700 + gq);
                float* hi = buf ? abuf1 : abuf0;
                float* lo = hi + CP * MIDP;
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 m4 = apply_act_t<PRE>(make_float4(racc[pl * CP + c4 * 4], racc[pl * CP + c4 * 4 + 1], racc[pl * CP + c4 * 4 + 2],
                                                                   racc[pl * CP + c4 * 4 + 3]), a.slope);
                    const float4 h = make_float4(tf32_rna(m4.x), tf32_rna(m4.y), tf32_rna(m4.z), tf32_rna(m4.w));
                    const float4 l = make_float4(m4.x - h.x, m4.y - h.y, m4.z - h.z, m4.w - h.w);
                    *reinterpret_cast<float4*>(hi + (c4 * MIDP + row) * 4) = h;
                    *reinterpret_cast<float4*>(lo + (c4 * MIDP + row) * 4) = l;
                }
                fence_async_smem();
#pragma unroll
                for (int k = 0; k < NPROD / 128; ++k) mbar_arrive(&a_full[buf]);    // barrier counts NPROD arrivals
            }
#pragma unroll
            for (int i = 0; i < NCOL; ++i) racc[i] = 0.f;
            drain(n_g2);
        }
        // ---- epilogue: this thread owns output row `row` (time step j0 + row), NCOL channels
        const int t = j0 + row;
        if (t < a.Tout) {
            const int co_l = co_tile * NT + dg * NCOL;          // channel within the group
            if (a.bias) {
#pragma unroll
                for (int i = 0; i < NCOL / 4; ++i) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + g * a.Cout_g + co_l) + i);
                    racc[4 * i] += b4.x; racc[4 * i + 1] += b4.y; racc[4 * i + 2] += b4.z; racc[4 * i + 3] += b4.w;
                }
            }
            if (a.res) {
                const float* rp = a.res + (long long)b * a.res_bs + (long long)t * a.ldr + g * a.r_goff + co_l;
#pragma unroll
                for (int i = 0; i < NCOL / 4; ++i) {
                    const float4 r4 = __ldg(reinterpret_cast<const float4*>(rp) + i);
                    racc[4 * i] = r4.x + racc[4 * i]; racc[4 * i + 1] = r4.y + racc[4 * i + 1];
                    racc[4 * i + 2] = r4.z + racc[4 * i + 2]; racc[4 * i + 3] = r4.w + racc[4 * i + 3];
                }
            }
            if (a.out_nct) {
                float* yp = a.y + (long long)b * a.y_bs + (long long)(g * a.y_goff + co_l) * a.Tout + t;
#pragma unroll
                for (int i = 0; i < NCOL; ++i) yp[(long long)i * a.Tout] = racc[i];
            } else {
                float* yp = a.y + (long long)b * a.y_bs + (long long)t * a.ldy + g * a.y_goff + co_l;
#pragma unroll
                for (int i = 0; i < NCOL / 4; ++i)
                    *(reinterpret_cast<float4*>(yp) + i) = make_float4(racc[4 * i], racc[4 * i + 1], racc[4 * i + 2], racc[4 * i + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();      // nobody exits while a peer can still multicast into / arrive on its smem
#ifdef ADEC_TIMELINE
    if (tid == 0 && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0) {
        const int ng = n_g1 + n_g2 < 64 ? n_g1 + n_g2 : 64;
        printf("TIMELINE NT=%d fuse=%d groups=%d pieces=%d end=%u\n", NT, (int)FUSE, n_g1 + n_g2, a.n_pieces, (unsigned)(clock64() - tl0_));
        for (int i = 0; i < ng; ++i)
            printf(" g%02d tma %6u | mma ready %6u issued %6u | drain got %6u done %6u\n", i, tl_[0][i], tl_[1][i], tl_[2][i], tl_[4][i], tl_[5][i]);
        for (int i = 0; i < a.n_pieces && i < 64; ++i) printf(" piece %d produced %6u\n", i, tl_[3][i]);
    }
#endif
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

}  // namespace adec
