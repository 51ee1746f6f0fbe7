// tc_conv_persist_kernel: persistent, cross-tile pipelined version of tc_conv_kernel (tc_kernels.cuh).
//
// Same math, same operand layouts, same grouped 3xTF32 accumulation.  What changes is the schedule: one CTA per SM
// loops over output tiles, and every role keeps running counters so its mbarrier rings simply continue from one tile
// into the next:
//     producers   : run ahead on the next tile's activation window while the current tile is in its 1x1 conv / epilogue
//     TMA         : keeps the weight ring full across tile boundaries
//     MMA (2-3 warps, groups round robin): GEMM 1 of tile i+1 starts as soon as its window piece and a TMEM partial are free;
//                   at NT=32 it is even issued BEFORE the 1x1 conv of tile i (PIPE)
//     drain       : 3-8 TMEM partials (instead of two) let the MMAs run ahead while these warps convert the fused
//                   intermediate or write the epilogue
// Barrier rule: an mbarrier parity wait cannot tell phase k from phase k-2, so every barrier here has
// waiters that observe each of its phases in order and can never fall two phases behind:
//     b_full[s] / p_empty[pb] : stage s = c % STAGES and partial pb = c % NPB always belong to the same issuer warp (NW divides both)
//     m_empty[g]              : "intermediate buffer free" is signalled to the drain group that writes that buffer NEXT
//     w_full/w_empty, m_full, p_full, b_empty : every waiter waits for every phase, and a phase cannot complete before all of
//                               them consumed the previous one
#pragma once
#include "tc_kernels.cuh"

namespace adec {

template <int NT> struct TcpCfg {
    static constexpr int MB = NT == 64 ? 2 : 1;            // fused-intermediate smem buffers (smem permitting; 2 stages + 2 buffers measured slower at NT=128,
                                                           // 6 stages + 1 buffer slower at NT=64)
    static constexpr int STAGES = TcCfg<NT>::STAGES;       // weight stages (at NT=32: 4 stages measured 13 % SLOWER than 3, 8 no faster)
    // MMA issuer warps take groups c = 0,1,2,... round robin.  Every mbarrier must have waiters that see each of its phases in
    // order (a parity wait cannot tell phase k from phase k-2), so a weight stage (c % STAGES) and a TMEM partial (c % NPB) must
    // always belong to the same warp: NW divides both.  Odd ring depth -> three issuer warps.
    static constexpr int NW = (STAGES % 2) ? 3 : 2;
    static constexpr int NPB = NT == 128 ? (NW == 3 ? 3 : 4) : (NW == 3 ? 6 : 8);   // TMEM partial buffers (NPB*NT <= 512 columns): how far the MMAs run ahead
    static constexpr int NDG = NT == 32 ? 2 : TcCfg<NT>::NDG;          // drain groups; at NT=32 each owns HALF a 32-column piece
    static constexpr int THREADS = 128 + TcCfg<NT>::NPROD + 128 * NDG;
};

template <int NT, bool FUSE, int PRE>
__global__ void __launch_bounds__(TcpCfg<NT>::THREADS, 1) tc_conv_persist_kernel(const ConvArgs a, int n_xtiles, int n_ytiles, int n_tiles) {
    using Cfg = TcCfg<NT>;
    constexpr int S = TcpCfg<NT>::STAGES, BST = Cfg::B_STAGE_FLOATS, CP = TC_CP, TT = TC_TT, NDG = TcpCfg<NT>::NDG;
    constexpr int NPROD = Cfg::NPROD, DRAIN0 = Cfg::DRAIN0, MIDP = TC_MIDP, NPB = TcpCfg<NT>::NPB;
    constexpr int NCOL = NT / NDG;                       // accumulator registers per drain thread
    constexpr bool HALF = NCOL < CP;                     // NT=32: a drain group owns 16 of the piece's 32 columns
    constexpr int PPG = HALF ? 1 : NCOL / CP;            // 32-column pieces (or half pieces) owned by one drain group
    constexpr int UC = HALF ? NCOL : CP;                 // columns per owned unit
    constexpr int MB = TcpCfg<NT>::MB;
    constexpr bool PIPE = FUSE && NT == 32;              // conv of tile i+1 issued before the 1x1 conv of tile i (needs 2*NCOL drain registers;
                                                         // measured: -19 % at NT=32, +3 % (register spills) at NT=64)
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr int NW = TcpCfg<NT>::NW;
    static_assert(S % NW == 0 && NPB % NW == 0, "a weight stage / TMEM partial must belong to one MMA warp");
    constexpr uint32_t TMEM_COLS = NPB * NT <= 32 ? 32 : NPB * NT <= 64 ? 64 : NPB * NT <= 128 ? 128 : NPB * NT <= 256 ? 256 : 512;   // power of two
    static_assert(NPB * NT <= 512, "TMEM has 512 columns");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* b_full = reinterpret_cast<uint64_t*>(smem_raw);     // [S]   weights landed
    uint64_t* b_empty = b_full + S;                                // [S]   weights consumed
    uint64_t* w_full = b_empty + S;                                // [2]   window piece written
    uint64_t* w_empty = w_full + 2;                                // [2]   window piece consumed
    uint64_t* m_full = w_empty + 2;                                // [MB]  fused-intermediate piece written
    uint64_t* m_empty = m_full + 2;                                // [2]   ... consumed; indexed by the drain group that writes the freed buffer NEXT
    uint64_t* p_full = m_empty + 2;                                // [NPB] TMEM partial complete
    uint64_t* p_empty = p_full + NPB;                              // [NPB] TMEM partial drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_empty + NPB);
    float* bst = reinterpret_cast<float*>(smem_raw + 512);        // up to 40 barriers + the TMEM slot live in the first 512 B
    const int wrows = TT + (a.Ktaps - 1) * a.dil;
    const int wrp = (wrows > MIDP ? wrows : MIDP) | 1;             // odd row pitch: conflict-free producer stores
    float* wbuf0 = bst + S * BST;
    float* wbuf1 = wbuf0 + 2 * CP * wrp;
    float* mbuf = wbuf1 + 2 * CP * wrp;                            // FUSE only: MB x 2*CP*MIDP floats

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int n_g1 = a.n_pieces * a.Ktaps;
    const int n_g2 = FUSE ? NT / CP : 0;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&w_full[i], NPROD); mbar_init(&w_empty[i], NW); }
        for (int i = 0; i < MB; ++i) mbar_init(&m_full[i], HALF ? 256 : 128);
        for (int i = 0; i < 2; ++i) mbar_init(&m_empty[i], NW);
        for (int i = 0; i < NPB; ++i) { mbar_init(&p_full[i], 1); mbar_init(&p_empty[i], 128 * NDG); }
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
#ifdef ADEC_TIMELINE
    __shared__ unsigned tlp_[6][64];
    __shared__ unsigned tlt_[8];
    const long long tlp0_ = clock64();
#define TLP(role, i) do { if (ti == 2 && (i) < 64) tlp_[role][(i)] = (unsigned)(clock64() - tlp0_); } while (0)
#else
#define TLP(role, i) do { } while (0)
#endif

    if (warp == 0) {
        // ------------------------------------------------ weight producer
        if (lane == 0) {
            int c = 0;
            auto stream = [&](const float* base, int n) {
                for (int k = 0; k < n; ++k, ++c) {
                    const int s = c % S, it = c / S;
                    if (it > 0) mbar_wait(&b_empty[s], (it - 1) & 1, 100);
                    mbar_arrive_expect_tx(&b_full[s], BST * 4);
                    bulk_g2s(bst + s * BST, base + (long long)k * BST, BST * 4, &b_full[s]);
                }
            };
            auto w1_of = [&](int tile) { return a.w + (long long)((tile / n_xtiles) % n_ytiles) * a.w_tile_floats; };
            if (PIPE) {
                // same order as the MMA warps: G1(t0), then per tile { G1(next), G2(this) }
                if ((int)blockIdx.x < n_tiles) stream(w1_of(blockIdx.x), n_g1);
                for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                    if (tile + (int)gridDim.x < n_tiles) stream(w1_of(tile + gridDim.x), n_g1);
                    stream(a.w2, n_g2);
                }
            } else {
                for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                    stream(w1_of(tile), n_g1);
                    if (FUSE) stream(a.w2, n_g2);
                }
            }
        }
    } else if (warp >= 1 && warp <= NW) {
        // ------------------------------------------------ MMA issuers (groups round robin)
        const int mw = warp - 1;
        int c = 0, wp = 0, mp = 0;
        int ti = -1, c_tile0 = 0;
        const uint32_t b_lbo = (uint32_t)NT * 16u;
        const uint32_t wbuf0_u = smem_u32(wbuf0), wbuf1_u = smem_u32(wbuf1), mbuf_u = smem_u32(mbuf), bst_u = smem_u32(bst);
        auto issue_group = [&](uint32_t a_hi, uint32_t a_lo, uint32_t lbo, uint32_t row_off) {
            const int s = c % S, pb = c % NPB;
            mbar_wait(&b_full[s], (c / S) & 1, 300);
            if (c >= NPB) mbar_wait(&p_empty[pb], ((c / NPB) - 1) & 1, 400);
            tc_fence_after();
            if (lane == 0) TLP(1, c - c_tile0);
            const uint32_t b_hi = bst_u + (uint32_t)s * (BST * 4u);
            const uint32_t b_lo = b_hi + (uint32_t)CP * NT * 4u;
            const uint32_t acc = tmem + (uint32_t)pb * NT;
            if (elect_one()) {
#pragma unroll
                for (int k8 = 0; k8 < CP / 8; ++k8)
                    umma_tf32(acc, umma_desc(a_lo + (uint32_t)(k8 * 2) * lbo + row_off, lbo), umma_desc(b_hi + (uint32_t)(k8 * 2) * b_lbo, b_lbo),
                              IDESC, k8 ? 1u : 0u);
#pragma unroll
                for (int k8 = 0; k8 < CP / 8; ++k8)
                    umma_tf32(acc, umma_desc(a_hi + (uint32_t)(k8 * 2) * lbo + row_off, lbo), umma_desc(b_lo + (uint32_t)(k8 * 2) * b_lbo, b_lbo),
                              IDESC, 1u);
#pragma unroll
                for (int k8 = 0; k8 < CP / 8; ++k8)
                    umma_tf32(acc, umma_desc(a_hi + (uint32_t)(k8 * 2) * lbo + row_off, lbo), umma_desc(b_hi + (uint32_t)(k8 * 2) * b_lbo, b_lbo),
                              IDESC, 1u);
                umma_commit(&b_empty[s]);
                umma_commit(&p_full[pb]);
                TLP(2, c - c_tile0);
            }
            __syncwarp();
        };
        const uint32_t lbo1 = (uint32_t)wrp * 16u, lbo2 = (uint32_t)MIDP * 16u;
        auto gemm1 = [&]() {        // one tile's conv over its window pieces
            for (int p = 0; p < a.n_pieces; ++p, ++wp) {
                const int buf = wp & 1;
                mbar_wait(&w_full[buf], (wp >> 1) & 1, 200);
                const uint32_t a_hi = buf ? wbuf1_u : wbuf0_u;
                const uint32_t a_lo = a_hi + (uint32_t)(CP / 4) * lbo1;
                for (int tap = 0; tap < a.Ktaps; ++tap, ++c)
                    if (c % NW == mw) issue_group(a_hi, a_lo, lbo1, (uint32_t)(tap * a.dil) * 16u);
                if (elect_one()) umma_commit(&w_empty[buf]);
                __syncwarp();
            }
        };
        auto gemm2 = [&]() {        // the fused 1x1 conv over the intermediate pieces
            for (int p = 0; p < NT / CP; ++p, ++mp, ++c) {
                const int mb = mp % MB;
                mbar_wait(&m_full[mb], (mp / MB) & 1, 250);
                const uint32_t m_hi = mbuf_u + (uint32_t)mb * (2u * CP * MIDP * 4u);
                if (c % NW == mw) issue_group(m_hi, m_hi + (uint32_t)(CP / 4) * lbo2, lbo2, 0u);
                // buffer mb is free for piece mp + MB, which drain group (mp + MB) % NDG writes: signal THAT group's barrier, so
                // that every m_empty barrier has one group of waiters which sees each of its phases exactly once, in order
                if (elect_one()) umma_commit(&m_empty[HALF ? 0 : (mp + MB) % NDG]);
                __syncwarp();
            }
        };
        if (PIPE) {
            // software-pipelined order: the conv of tile i+1 is issued BEFORE the 1x1 conv of tile i, so the tensor pipe
            // works on it while the drain warps turn tile i's accumulators into the 1x1 conv's operand
            if ((int)blockIdx.x < n_tiles) { ++ti; c_tile0 = c; gemm1(); }
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                if (tile + (int)gridDim.x < n_tiles) { ++ti; c_tile0 = c; gemm1(); }
                gemm2();
            }
        } else {
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                ++ti; c_tile0 = c;
                gemm1();
                if (FUSE) gemm2();
            }
        }
    } else if (warp >= 4 && warp < DRAIN0) {
        // ------------------------------------------------ activation producers
        const int pt = tid - 128;
        int wp = 0;
        constexpr int RPP = NPROD / 8;
        constexpr int UNR = 6;
        const int c4 = pt & 7, m0 = pt >> 3;
        int ti = -1;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            ++ti;
            const int xt = tile % n_xtiles, y = (tile / n_xtiles) % n_ytiles, b = tile / (n_xtiles * n_ytiles);
            const int j0 = xt * TT, g = y / a.n_co_tiles, co_tile = y - g * a.n_co_tiles;
            const float* xg = a.x + (long long)b * a.x_bs + g * a.x_goff;
            const float* sg = a.st_in + (long long)b * a.P * a.st_ld + g * a.st_goff;
            for (int p = 0; p < a.n_pieces; ++p, ++wp) {
                const int buf = wp & 1;
                if (wp >= 2) mbar_wait(&w_empty[buf], ((wp >> 1) - 1) & 1, 500);
                float* hi = buf ? wbuf1 : wbuf0;
                float* lo = hi + CP * wrp;
                const int q = p * CP + c4 * 4;
                int r = 0, ci = q;
                if (a.RG > 1) { r = q >> a.lgCin; ci = q & (a.Cin - 1); }
                const float* srow = sg + ci;
                const float* xrow = xg + ci;
                float* hcol = hi + (c4 * wrp) * 4;
                float* lcol = lo + (c4 * wrp) * 4;
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
                                if ((i >= a.P || a.hist_rep) && i - a.P < a.T) {
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
                                const float4 l = make_float4(x4.x - h.x, x4.y - h.y, x4.z - h.z, x4.w - h.w);
                                *reinterpret_cast<float4*>(hcol + m * 4) = h;
                                *reinterpret_cast<float4*>(lcol + m * 4) = l;
                            }
                        }
                    }
                }
                fence_async_smem();
                mbar_arrive(&w_full[buf]);
                if (pt == 0) TLP(3, p);
            }
            // ---- new causal state (conv_layer.py:155)
            if (xt == (a.Tout - 1) / TT && co_tile == 0 && g < a.st_groups && a.P > 0) {
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
        }
    } else if (warp >= DRAIN0) {
        // ------------------------------------------------ drain warps: register accumulation, fused intermediate, epilogue
        const int dg = (warp - DRAIN0) >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        // Column ownership: piece pl of group dg is 32-column piece (pl*NDG + dg), so that consecutive intermediate
        // pieces alternate between the groups and every group meets every m_empty phase in order.
        float racc[NCOL];
        constexpr bool PREFETCH_RES = FUSE && NT <= 64;     // registers permitting
        float4 rpre[PREFETCH_RES ? PPG : 1][UC / 4];
        int c = 0, mq = 0, ti = -1, c_tile0 = 0;
        int m_waits = 0;
        auto drain = [&](float (&acc)[NCOL], int ngroups) {
            for (int gi = 0; gi < ngroups; ++gi, ++c) {
                const int pb = c % NPB;
                mbar_wait(&p_full[pb], (c / NPB) & 1, 600);
                tc_fence_after();
                if (tid == DRAIN0 * 32) TLP(4, c - c_tile0);
                if (HALF) {
                    uint32_t r0[16];
                    tmem_ld16(tmem + lane_base + (uint32_t)pb * NT + (uint32_t)dg * UC, r0);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[i] = __fadd_rn(acc[i], __uint_as_float(r0[i]));
                } else {
#pragma unroll
                    for (int pl = 0; pl < PPG; ++pl) {
                        const uint32_t taddr = tmem + lane_base + (uint32_t)pb * NT + (uint32_t)(pl * NDG + dg) * CP;
                        uint32_t r0[16], r1[16];
                        tmem_ld16(taddr, r0);
                        tmem_ld16(taddr + 16, r1);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            acc[pl * UC + i] = __fadd_rn(acc[pl * UC + i], __uint_as_float(r0[i]));
                            acc[pl * UC + 16 + i] = __fadd_rn(acc[pl * UC + 16 + i], __uint_as_float(r1[i]));
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&p_empty[pb]);
                if (tid == DRAIN0 * 32) TLP(5, c - c_tile0);
            }
        };
        float oacc[PIPE ? NCOL : 1];                         // PIPE: 1x1-conv sums of tile i while racc already holds tile i+1
        if (PIPE && (int)blockIdx.x < n_tiles) {
#pragma unroll
            for (int i = 0; i < NCOL; ++i) racc[i] = 0.f;
            drain(racc, n_g1);                                // conv of the first tile
        }
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            ++ti; c_tile0 = c;
#ifdef ADEC_TIMELINE
            if (tid == DRAIN0 * 32 && ti >= 1 && ti <= 4) tlt_[ti] = (unsigned)(clock64() - tlp0_);
#endif
            const int xt = tile % n_xtiles, y = (tile / n_xtiles) % n_ytiles, b = tile / (n_xtiles * n_ytiles);
            const int j0 = xt * TT, g = y / a.n_co_tiles, co_tile = y - g * a.n_co_tiles;
            if (!PIPE) {
#pragma unroll
                for (int i = 0; i < NCOL; ++i) racc[i] = 0.f;
                drain(racc, n_g1);
            }
            if (FUSE) {
                // activation first (registers only), so that it overlaps the wait for a free intermediate buffer
#pragma unroll
                for (int i = 0; i < NCOL; i += 4) {
                    const float4 m4 = apply_act_t<PRE>(make_float4(racc[i], racc[i + 1], racc[i + 2], racc[i + 3]), a.slope);
                    racc[i] = m4.x; racc[i + 1] = m4.y; racc[i + 2] = m4.z; racc[i + 3] = m4.w;
                }
                // intermediate pieces in consumption order 0,1,2,...: piece q is written by group q % NDG (both groups at NT=32)
                // into buffer Q % MB (Q = running piece index), which is free once piece Q-MB was consumed; the MMA warps signal
                // that on the writing group's own m_empty barrier, so the k-th wait of a thread is for that barrier's k-th phase.
#pragma unroll
                for (int q = 0; q < NT / CP; ++q) {
                    const int Q = mq + q, mb = Q % MB;
                    if (HALF || q % NDG == dg) {
                        if (Q >= MB) { mbar_wait(&m_empty[HALF ? 0 : dg], m_waits & 1, 700); ++m_waits; }
                        const int pl = HALF ? 0 : q / NDG;
                        float* hi = mbuf + mb * (2 * CP * MIDP);
                        float* lo = hi + CP * MIDP;
#pragma unroll
                        for (int u = 0; u < UC / 4; ++u) {
                            const int c4 = HALF ? dg * (UC / 4) + u : u;       // 16-byte channel column inside the 32-channel piece
                            const float4 m4 = make_float4(racc[pl * UC + u * 4], racc[pl * UC + u * 4 + 1], racc[pl * UC + u * 4 + 2],
                                                          racc[pl * UC + u * 4 + 3]);
                            const float4 h = make_float4(tf32_rna(m4.x), tf32_rna(m4.y), tf32_rna(m4.z), tf32_rna(m4.w));
                            const float4 l = make_float4(m4.x - h.x, m4.y - h.y, m4.z - h.z, m4.w - h.w);
                            *reinterpret_cast<float4*>(hi + (c4 * MIDP + row) * 4) = h;
                            *reinterpret_cast<float4*>(lo + (c4 * MIDP + row) * 4) = l;
                        }
                        fence_async_smem();
                        mbar_arrive(&m_full[mb]);
                    }
                }
                mq += NT / CP;
#pragma unroll
                for (int i = 0; i < NCOL; ++i) racc[i] = 0.f;
                if (PREFETCH_RES && a.res && j0 + row < a.Tout) {
                    // the skip tensor's rows are known now: fetch them while the MMAs run
#pragma unroll
                    for (int pl = 0; pl < PPG; ++pl) {
                        const int co_l = co_tile * NT + (HALF ? dg * UC : (pl * NDG + dg) * CP);
                        const float* rp = a.res + (long long)b * a.res_bs + (long long)(j0 + row) * a.ldr + g * a.r_goff + co_l;
#pragma unroll
                        for (int i = 0; i < UC / 4; ++i) rpre[pl][i] = __ldg(reinterpret_cast<const float4*>(rp) + i);
                    }
                }
                if (PIPE) {
                    // partials arrive in MMA issue order: the NEXT tile's conv first (into racc), then this tile's 1x1 conv
                    if (tile + (int)gridDim.x < n_tiles) drain(racc, n_g1);
#pragma unroll
                    for (int i = 0; i < (PIPE ? NCOL : 1); ++i) oacc[i] = 0.f;
                    drain(reinterpret_cast<float (&)[NCOL]>(oacc), n_g2);
                } else {
                    drain(racc, n_g2);
                }
            }
            float* const outv = PIPE ? oacc : racc;
            // ---- epilogue: row `row` of the tile, this group's PPG pieces of 32 channels
            const int t = j0 + row;
            if (t < a.Tout) {
#pragma unroll
                for (int pl = 0; pl < PPG; ++pl) {
                    const int co_l = co_tile * NT + (HALF ? dg * UC : (pl * NDG + dg) * CP);
                    if (co_l >= a.Cout_g) continue;            // zero-padded part of a channel tile (e.g. 96 outputs in a 128-wide tile)
                    float* v = outv + pl * UC;
                    if (a.bias) {
#pragma unroll
                        for (int i = 0; i < UC / 4; ++i) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + g * a.Cout_g + co_l) + i);
                            v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
                        }
                    }
                    if (a.res) {
                        const float* rp = a.res + (long long)b * a.res_bs + (long long)t * a.ldr + g * a.r_goff + co_l;
                        float4 r4[UC / 4];
#pragma unroll
                        for (int i = 0; i < UC / 4; ++i)
                            r4[i] = PREFETCH_RES ? rpre[PREFETCH_RES ? pl : 0][i] : __ldg(reinterpret_cast<const float4*>(rp) + i);   // all loads in flight first
#pragma unroll
                        for (int i = 0; i < UC / 4; ++i) {
                            v[4 * i] = r4[i].x + v[4 * i]; v[4 * i + 1] = r4[i].y + v[4 * i + 1];
                            v[4 * i + 2] = r4[i].z + v[4 * i + 2]; v[4 * i + 3] = r4[i].w + v[4 * i + 3];
                        }
                    }
                    if (a.out_nct) {
                        float* yp = a.y + (long long)b * a.y_bs + (long long)(g * a.y_goff + co_l) * a.Tout + t;
#pragma unroll
                        for (int i = 0; i < UC; ++i) yp[(long long)i * a.Tout] = v[i];
                    } else {
                        float* yp = a.y + (long long)b * a.y_bs + (long long)t * a.ldy + g * a.y_goff + co_l;
#pragma unroll
                        for (int i = 0; i < UC / 4; ++i)
                            *(reinterpret_cast<float4*>(yp) + i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
#ifdef ADEC_TIMELINE
    if (tid == 0 && blockIdx.x == 1) {
        const int ng = n_g1 + n_g2 < 64 ? n_g1 + n_g2 : 64;
        printf("PTIMELINE NT=%d fuse=%d groups=%d pieces=%d tile starts(drain) t1=%u t2=%u t3=%u t4=%u end=%u\n", NT, (int)FUSE, n_g1 + n_g2,
               a.n_pieces, tlt_[1], tlt_[2], tlt_[3], tlt_[4], (unsigned)(clock64() - tlp0_));
        for (int i = 0; i < ng; ++i)
            printf(" g%02d mma ready %7u issued %7u | drain got %7u done %7u\n", i, tlp_[1][i], tlp_[2][i], tlp_[4][i], tlp_[5][i]);
        for (int i = 0; i < a.n_pieces && i < 64; ++i) printf(" piece %d produced %7u\n", i, tlp_[3][i]);
    }
#endif
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

}  // namespace adec
