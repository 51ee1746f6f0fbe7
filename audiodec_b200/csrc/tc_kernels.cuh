// tc_conv_kernel: the causal conv family on 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same ConvArgs contract as conv_gemm_kernel (kernels.cuh); different engine:
//   * GEMM orientation: M = 128 time steps (TMEM lanes), N = output channels (TMEM columns), K = input
//     channels of one tap.  A (activations) and B (weights) are both K-major, no-swizzle UMMA operands
//     stored as "column blocks"  smem[(k/4)*ROWS + row][k%4]  (16-byte core-matrix rows, SBO = 128 B,
//     LBO = ROWS*16 B).  In this layout a conv tap is a *row-shifted start address* of the same
//     window - no im2col, no copies: tap k of a dilated conv reads rows [k*dil, k*dil+128).
//   * precision: 3xTF32 error-compensated products (a = a_hi + a_lo, w = w_hi + w_lo in tf32;
//     acc += a_lo*w_hi + a_hi*w_lo + a_hi*w_hi, fp32 accumulation in TMEM).  Measured max error
//     ~1e-7 relative (tools/tc_probe.cu) - fp32-grade, which the bit-identical-index contract needs;
//     plain TF32 (2e-3) is not.
//   * warp roles: warp 0 = weight producer (cp.async.bulk / TMA 1-D, host-pre-split hi|lo tiles already in
//     UMMA layout) + TMEM allocator; warp 1 = MMA issuer (one thread); warps 2-5 = activation producers
//     (global -> pre-activation -> hi/lo split -> smem, 32-channel pieces, double buffered) and then the
//     epilogue (tcgen05.ld, bias / residual, stores).  mbarrier rings connect them; tcgen05.commit
//     releases smem back to the producers.
//   * FUSE: the residual unit keeps acc1 in TMEM, the producers turn it into the activated hi/lo operand of
//     the 1x1 conv piece by piece, acc2 is a second TMEM region; the skip tensor is added in the epilogue.
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

constexpr int TC_TT = 128;        // output rows per CTA (UMMA M)
constexpr int TC_CP = 32;         // channels per activation piece
constexpr int TC_THREADS = 192;   // warp0 TMA+alloc, warp1 MMA, warps 2-5 producers/epilogue
constexpr int TC_NPROD = 128;

template <int NT, int KS>
struct TcCfg {
    static constexpr int STAGES = NT == 256 ? 3 : (NT == 128 ? 3 : 4);
    static constexpr int B_STAGE_FLOATS = 2 * KS * NT;                      // hi | lo
};

template <int NT, int KS, bool FUSE>
__global__ void __launch_bounds__(TC_THREADS) tc_conv_kernel(const ConvArgs a) {
    using Cfg = TcCfg<NT, KS>;
    constexpr int S = Cfg::STAGES, BST = Cfg::B_STAGE_FLOATS, CP = TC_CP, TT = TC_TT;
    constexpr int KS_PER_PIECE = CP / KS;
    static_assert(CP % KS == 0 && KS % 8 == 0, "bad KS");
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t TMEM_COLS = FUSE ? 2 * NT : NT;
    constexpr int MIDP = 129;   // row pitch (rows) of the 1x1 conv's activated operand

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* b_full = reinterpret_cast<uint64_t*>(smem_raw);     // [S]
    uint64_t* b_empty = b_full + S;                                // [S]
    uint64_t* a_full = b_empty + S;                                // [2]
    uint64_t* a_empty = a_full + 2;                                // [2]
    uint64_t* acc_full = a_empty + 2;                              // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 2);
    float* bst = reinterpret_cast<float*>(smem_raw + 256);
    const int wrows = TT + (a.Ktaps - 1) * a.dil;
    const int wrp = (wrows > MIDP ? wrows : MIDP) | 1;             // odd row pitch: conflict-free producer stores
    float* abuf0 = bst + S * BST;
    float* abuf1 = abuf0 + 2 * CP * wrp;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int j0 = blockIdx.x * TT;
    const int g = blockIdx.y / a.n_co_tiles;
    const int co_tile = blockIdx.y - g * a.n_co_tiles;
    const int b = blockIdx.z;
    const int n_k1 = a.n_pieces * a.Ktaps * KS_PER_PIECE;
    const int n_k2 = FUSE ? (NT / CP) * KS_PER_PIECE : 0;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], TC_NPROD); mbar_init(&a_empty[i], 1); mbar_init(&acc_full[i], 1); }
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

    if (warp == 0) {
        // ------------------------------------------------ weight producer (TMA 1-D bulk copies)
        if (lane == 0) {
            const float* w1 = a.w + (long long)blockIdx.y * a.w_tile_floats;
            for (int c = 0; c < n_k1 + n_k2; ++c) {
                const int s = c % S, it = c / S;
                if (it > 0) mbar_wait(&b_empty[s], (it - 1) & 1);
                const float* src = c < n_k1 ? w1 + (long long)c * BST : a.w2 + (long long)(c - n_k1) * BST;
                mbar_arrive_expect_tx(&b_full[s], BST * 4);
                bulk_g2s(bst + s * BST, src, BST * 4, &b_full[s]);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            int c = 0, gp = 0;
            const uint32_t a_lbo = (uint32_t)wrp * 16u, b_lbo = (uint32_t)NT * 16u;
            for (int phase = 0; phase < (FUSE ? 2 : 1); ++phase) {
                const int pieces = phase == 0 ? a.n_pieces : NT / CP;
                const int taps = phase == 0 ? a.Ktaps : 1;
                const uint32_t lbo = phase == 0 ? a_lbo : (uint32_t)MIDP * 16u;
                const uint32_t acc = tmem + (phase == 0 ? 0 : NT);
                uint32_t accumulate = 0;
                for (int p = 0; p < pieces; ++p, ++gp) {
                    const int buf = gp & 1;
                    mbar_wait(&a_full[buf], (gp >> 1) & 1);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(buf ? abuf1 : abuf0);
                    const uint32_t a_lo = a_hi + (uint32_t)(CP / 4) * lbo;     // lo block follows the hi block
                    for (int tap = 0; tap < taps; ++tap) {
                        const uint32_t row_off = (uint32_t)(tap * a.dil) * 16u;
                        for (int ks = 0; ks < KS_PER_PIECE; ++ks, ++c) {
                            const int s = c % S;
                            mbar_wait(&b_full[s], (c / S) & 1);
                            tc_fence_after();
                            const uint32_t b_hi = smem_u32(bst + s * BST);
                            const uint32_t b_lo = b_hi + (uint32_t)KS * NT * 4u;
#pragma unroll
                            for (int k8 = 0; k8 < KS / 8; ++k8) {
                                const uint32_t ao = (uint32_t)((ks * KS) / 4 + k8 * 2) * lbo + row_off;
                                const uint32_t bo = (uint32_t)(k8 * 2) * b_lbo;
                                const uint64_t dah = umma_desc(a_hi + ao, lbo), dal = umma_desc(a_lo + ao, lbo);
                                const uint64_t dbh = umma_desc(b_hi + bo, b_lbo), dbl = umma_desc(b_lo + bo, b_lbo);
                                umma_tf32(acc, dal, dbh, IDESC, accumulate);   // small terms first
                                umma_tf32(acc, dah, dbl, IDESC, 1u);
                                umma_tf32(acc, dah, dbh, IDESC, 1u);
                                accumulate = 1u;
                            }
                            umma_commit(&b_empty[s]);       // weights stage free once these MMAs retire
                        }
                    }
                    umma_commit(&a_empty[buf]);             // activation piece free
                }
                umma_commit(&acc_full[phase]);              // accumulator complete
            }
        }
    } else {
        // ------------------------------------------------ activation producers, then epilogue (128 threads)
        const int pt = tid - 64;                            // 0..127
        const int row = (warp & 3) * 32 + lane;             // TMEM lane == output row this thread owns
        const float* xg = a.x + (long long)b * a.x_bs + g * a.x_goff;
        const float* sg = a.st_in + (long long)b * a.P * a.st_ld + g * a.st_goff;
        int gp = 0;
        for (int p = 0; p < a.n_pieces; ++p, ++gp) {
            const int buf = gp & 1;
            if (gp >= 2) mbar_wait(&a_empty[buf], ((gp >> 1) - 1) & 1);
            float* hi = buf ? abuf1 : abuf0;
            float* lo = hi + CP * wrp;
            const int nvec = wrows * (CP / 4);
            for (int idx = pt; idx < nvec; idx += TC_NPROD) {
                const int m = idx >> 3, c4 = idx & 7;
                const int q = p * CP + c4 * 4;
                int r = 0, ci = q;
                if (a.RG > 1) { r = q >> a.lgCin; ci = q & (a.Cin - 1); }
                const long long i = (long long)(j0 + m) * a.RG + r;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < a.P) {
                    v = *reinterpret_cast<const float4*>(sg + i * a.st_ld + ci);
                } else {
                    const long long t = i - a.P;
                    if (t < a.T) {
                        v = __ldg(reinterpret_cast<const float4*>(xg + t * a.ldx + ci));
                        if (a.pre_act == ACT_NORM) {
                            const float4 mu = *reinterpret_cast<const float4*>(a.mean + ci);
                            const float4 sc = *reinterpret_cast<const float4*>(a.scale + ci);
                            v.x = __fdiv_rn(v.x - mu.x, sc.x); v.y = __fdiv_rn(v.y - mu.y, sc.y);
                            v.z = __fdiv_rn(v.z - mu.z, sc.z); v.w = __fdiv_rn(v.w - mu.w, sc.w);
                        } else {
                            v = apply_act(v, a.pre_act, a.slope);
                        }
                    }
                }
                const float4 h = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
                const float4 l = make_float4(tf32_rna(v.x - h.x), tf32_rna(v.y - h.y), tf32_rna(v.z - h.z), tf32_rna(v.w - h.w));
                *reinterpret_cast<float4*>(hi + (c4 * wrp + m) * 4) = h;
                *reinterpret_cast<float4*>(lo + (c4 * wrp + m) * 4) = l;
            }
            fence_async_smem();
            mbar_arrive(&a_full[buf]);
        }
        if (FUSE) {
            // mid = act(acc1) -> hi/lo operand of the 1x1 conv, 32 columns (one piece) at a time
            mbar_wait(&acc_full[0], 0);
            tc_fence_after();
            for (int p = 0; p < NT / CP; ++p, ++gp) {
                const int buf = gp & 1;
                if (gp >= 2) mbar_wait(&a_empty[buf], ((gp >> 1) - 1) & 1);
                float* hi = buf ? abuf1 : abuf0;
                float* lo = hi + CP * MIDP;
                float v[32];
                tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + p * CP, v);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    float4 m4 = apply_act(make_float4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]), a.mid_act, a.slope);
                    const float4 h = make_float4(tf32_rna(m4.x), tf32_rna(m4.y), tf32_rna(m4.z), tf32_rna(m4.w));
                    const float4 l = make_float4(tf32_rna(m4.x - h.x), tf32_rna(m4.y - h.y), tf32_rna(m4.z - h.z), tf32_rna(m4.w - h.w));
                    *reinterpret_cast<float4*>(hi + (c4 * MIDP + row) * 4) = h;
                    *reinterpret_cast<float4*>(lo + (c4 * MIDP + row) * 4) = l;
                }
                tc_fence_before();
                fence_async_smem();
                mbar_arrive(&a_full[buf]);
            }
        }
        mbar_wait(&acc_full[FUSE ? 1 : 0], 0);
        tc_fence_after();
        // ---- epilogue: this thread owns output row `row` (time step j0 + row), all NT channels of the tile
        const int t = j0 + row;
        const uint32_t tacc = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (FUSE ? NT : 0);
        for (int c0 = 0; c0 < NT; c0 += 32) {
            float v[32];
            tmem_ld32(tacc + c0, v);            // .sync.aligned: executed by the whole warp, also for rows past Tout
            if (t < a.Tout) {
                const int co_l = co_tile * NT + c0;
                if (a.bias) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += __ldg(a.bias + g * a.Cout_g + co_l + i);
                }
                if (a.res) {
                    const float* rp = a.res + (long long)b * a.res_bs + (long long)t * a.ldr + g * a.r_goff + co_l;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 r4 = __ldg(reinterpret_cast<const float4*>(rp) + i);
                        v[4 * i] = r4.x + v[4 * i]; v[4 * i + 1] = r4.y + v[4 * i + 1];
                        v[4 * i + 2] = r4.z + v[4 * i + 2]; v[4 * i + 3] = r4.w + v[4 * i + 3];
                    }
                }
                if (a.out_nct) {
                    float* yp = a.y + (long long)b * a.y_bs + (long long)(g * a.y_goff + co_l) * a.Tout + t;
#pragma unroll
                    for (int i = 0; i < 32; ++i) yp[(long long)i * a.Tout] = v[i];
                } else {
                    float* yp = a.y + (long long)b * a.y_bs + (long long)t * a.ldy + g * a.y_goff + co_l;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        *(reinterpret_cast<float4*>(yp) + i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
        }
        // ---- new causal state (conv_layer.py:155)
        if (blockIdx.x == gridDim.x - 1 && co_tile == 0 && g < a.st_groups && a.P > 0) {
            float* so = a.st_out + (long long)b * a.P * a.st_ld + g * a.st_goff;
            const int nvec = a.P * (a.Cin / 4);
            for (int idx = pt; idx < nvec; idx += TC_NPROD) {
                const int r = idx / (a.Cin / 4);
                const int ci = (idx - r * (a.Cin / 4)) * 4;
                const long long i = (long long)a.T + r;
                float4 v;
                if (i < a.P) {
                    v = *reinterpret_cast<const float4*>(sg + i * a.st_ld + ci);
                } else {
                    v = __ldg(reinterpret_cast<const float4*>(xg + (i - a.P) * a.ldx + ci));
                    if (a.pre_act == ACT_NORM) {
                        const float4 mu = *reinterpret_cast<const float4*>(a.mean + ci);
                        const float4 sc = *reinterpret_cast<const float4*>(a.scale + ci);
                        v.x = __fdiv_rn(v.x - mu.x, sc.x); v.y = __fdiv_rn(v.y - mu.y, sc.y);
                        v.z = __fdiv_rn(v.z - mu.z, sc.z); v.w = __fdiv_rn(v.w - mu.w, sc.w);
                    } else {
                        v = apply_act(v, a.pre_act, a.slope);
                    }
                }
                *reinterpret_cast<float4*>(so + (long long)r * a.st_ld + ci) = v;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

}  // namespace adec
