// audiodec_b200 device kernels (sm_100a).
//
// Everything the reference does with torch.cat + nn.Conv1d / nn.ConvTranspose1d + separate
// activation / add kernels (layers/conv_layer.py:153-156,194-197; residual_unit.py:78-81) is one
// kernel family here:
//
//   conv_gemm_kernel  - stateful causal conv as an im2col-free implicit GEMM on CUDA cores
//                       (fp32 FFMA: the 1e-4 / bit-identical-index contract rules out TF32/bf16).
//                       Activations are channels-last (B,T,C) in HBM so a time window is one
//                       contiguous range and every load is a 128-bit channel vector.  The CTA keeps
//                       its input window [t0-halo, t0+TT) x C in shared memory (pre-activation and
//                       causal history applied while it is written), weight tiles are streamed
//                       L2 -> smem by a producer warp with cp.async.bulk (TMA, 1-D) through a
//                       4-stage mbarrier ring, 8x8 register micro-tiles accumulate, and the epilogue
//                       fuses bias / residual / layout.  With FUSE the whole residual unit
//                       ELU -> k7 dilated -> ELU -> 1x1 -> +x runs without leaving the SM.
//   stem_kernel       - Cin = 1 first conv (pure store bandwidth).
//   head_kernel       - Cout = 1 last conv (+ LeakyReLU / bias / tanh for the vocoder).
//   rvq_kernel        - 8-stage residual VQ: fp32 distances in the reference's rounding order,
//                       first-index arg-min by warp shuffles, int64 flat indices.
//   lookup_kernel     - codebook gather-sum.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace adec {

enum PreAct { ACT_NONE = 0, ACT_ELU = 1, ACT_LRELU = 2, ACT_NORM = 3 };

// ------------------------------------------------------------------------------------------------
// small PTX helpers (mbarrier + 1-D bulk async copy)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint expires)
// instead of spinning - polling loops were 30-40 % of all issued instructions in the first tensor-core profile.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity), "r"(0x989680u)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must trap (kernel error) instead of hanging the GPU.  The clock is only consulted every
// 64 failed probes so the common path stays a 2-instruction loop.  The bound is ~20 s of SM clocks: far beyond any legitimate wait
// even when the context is time-sliced with other tenants (clock64 keeps counting while descheduled); -DADEC_NO_WATCHDOG compiles
// the check out (plain try_wait loop with the hardware suspend hint).
#ifndef ADEC_WATCHDOG_CYCLES
#define ADEC_WATCHDOG_CYCLES 40000000000LL
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity, int tag = 0) {
    if (mbar_try_wait(b, parity)) return;
#ifdef ADEC_NO_WATCHDOG
    while (!mbar_try_wait(b, parity)) { }
#else
    const long long t0 = clock64();
    while (true) {
        // 64 bare probes (2 instructions each: the polling loop was 16 % of all issued instructions with the clock test inside it)
#pragma unroll 1
        for (int it = 0; it < 64; ++it) {
            if (mbar_try_wait(b, parity)) return;
#ifdef ADEC_WAIT_SLEEP
            if (tag != 200 && tag != 250 && tag != 300 && tag != 400) __nanosleep(ADEC_WAIT_SLEEP);   // not the MMA issuers: they are the critical path
#endif
        }
        if (clock64() - t0 > ADEC_WATCHDOG_CYCLES) {
#ifdef ADEC_WATCHDOG_PRINT
            printf("adec: mbarrier wait timed out: tag %d parity %u block (%d,%d,%d) thread %d\n", tag, parity, blockIdx.x, blockIdx.y,
                   blockIdx.z, threadIdx.x);
#endif
            __trap();
        }
    }
#endif
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// nn.ELU(alpha=1): x > 0 ? x : expm1(x).  expm1f() costs ~40 instructions and the elementwise work, not the tensor
// pipe, bounds the narrow layers; ex2.approx(x*log2e) - 1 is 4 instructions and within 2.4e-7 ABSOLUTE of expm1 on
// (-inf, 0] (2^-22 relative error of ex2.approx on a value <= 1) - the same order as the fp32 rounding of the O(1)
// activations it is summed with.  Parity margins: tests/test_layers_gpu.py (1e-5 on a fused unit), golden indices.
// __expf() wraps ex2.approx in a denormal-range fix-up (FSETP -126 / FMUL 0.5 / FMUL square: 8 instructions per ELU, 18 % of all
// instructions of the C = 32 unit in the round-2 ncu capture); ex2.approx.ftz is the bare MUFU.EX2 and gives the same ELU bit for bit:
// where the two differ (x log2(e) < -126) exp(x) is below 2^-126 and exp(x) - 1 rounds to -1 either way.
__device__ __forceinline__ float act_elu(float v) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * 1.4426950408889634f));
    e -= 1.0f;
    return v > 0.f ? v : e;
}
__device__ __forceinline__ float act_lrelu(float v, float s) { return v > 0.f ? v : v * s; }

__device__ __forceinline__ float4 apply_act(float4 v, int act, float slope) {
    if (act == ACT_ELU) {
        v.x = act_elu(v.x); v.y = act_elu(v.y); v.z = act_elu(v.z); v.w = act_elu(v.w);
    } else if (act == ACT_LRELU) {
        v.x = act_lrelu(v.x, slope); v.y = act_lrelu(v.y, slope); v.z = act_lrelu(v.z, slope); v.w = act_lrelu(v.w, slope);
    }
    return v;
}

// ------------------------------------------------------------------------------------------------
// conv_gemm_kernel
// ------------------------------------------------------------------------------------------------
// The conv is evaluated as  Y[t][co] = sum_{piece,tap,ci} Xw[t + tap*dil][piece*CW + ci] * W[tap][ci][co]
// over an "extended" input x~ = history(P rows) || chunk(T rows) (layers/conv_layer.py:154).
//   * stride-s convs (k = 2s) use RG = s: s consecutive x~ rows are folded into one window row of
//     s*Cin channels, which turns them into 2-tap stride-1 convs (and keeps smem reads conflict-free);
//   * transposed convs (k = 2s, crop [s:-s], conv_layer.py:197) are 2-tap convs with s*Cout outputs:
//     y[j*s+r] = b + W[:,:,r]^T x[j] + W[:,:,s+r]^T x[j-1]; the (T, s*Cout) result *is* (T*s, Cout).
struct ConvArgs {
    // input activations, channels-last; group g reads channels [g*x_goff, g*x_goff + Cin)
    const float* x;
    long long x_bs;
    int ldx, x_goff;
    // causal state (history rows), (B, P, st_ld); ping-pong in/out
    const float* st_in;
    float* st_out;
    int st_ld, st_goff, st_groups;   // st_groups: how many groups own distinct state channels
    int P, T, Tout;
    int Ktaps, dil, RG, lgCin, Cin;  // Cin: channels per x~ row (per group)
    int n_pieces;
    int pre_act;
    float slope;
    const float* mean;   // ACT_NORM: (v - mean[c]) / scale[c]   (HiFiGAN.py:276-279)
    const float* scale;
    // weights (packed by the host, see pack_weights in adec.cu)
    const float* w;
    const float* w2;     // FUSE: the residual unit's 1x1 conv
    const float* bias;   // [g*Cout_g + co] or nullptr
    int n_co_tiles, Cout_g;
    long long w_tile_floats;   // floats per (group, co_tile) of w
    // residual (raw), output
    const float* res;
    long long res_bs;
    int ldr, r_goff;
    float* y;
    long long y_bs;
    int ldy, y_goff, out_nct;
    int mid_act;         // FUSE: activation between the two GEMMs
    int hist_rep;        // non-streaming forward of a transposed conv: history rows = the FIRST input row (ReplicationPad1d,
                         // conv_layer.py:189-192) instead of the stored state
    // tcgen05 kind::f16 engine (tc_f16.cuh): weights are stored times a power of two; the drain warps multiply the sums by these
    float w_scale, w2_scale;
    int n_wbuf;          // window buffers in shared memory (2..4)
    int teams;           // tc_f16, un-fused launches: 2 = the 8 producer warps build alternate window pieces as two teams of 4, 1 = one team
    int gspan;           // tc_f16: 1 = one TMEM partial per 32-channel piece (all taps) instead of one per tap pair
    // stacked rows: when a stream contributes fewer rows than a 128-row tile, the tiles run over ONE row space in which stream s owns
    // rows [s * stack_L, (s + 1) * stack_L), stack_L = Tout + (Ktaps - 1) * dil: local rows >= Tout are the receptive-field overlap into
    // the next stream and are computed but never stored.  0 = one row space per stream (blockIdx-style b dimension).
    int stack_L, n_streams;
    int* err;            // device flag word: bit 1 = an activation left the fp16-split range (|a| >= 6e4)
    int dbg_flags;       // timing experiments only (ADEC_DBG_FLAGS): bit 0 = producers skip loads / activation / split of interior pieces
    int dbg_wdiv;        // timing experiments only (ADEC_DBG_WDIV): 0/1 = normal, k > 1 = stream 1/k of every weight stage, -1 = none (results are WRONG)
    unsigned int* tl;    // -DADEC_TIMELINE builds: event buffer of one CTA ({code << 24 | index, clock} pairs; word 0 = count), nullptr = off
    unsigned long long* dbg;   // ADEC_KTRACE: {globaltimer at start, at end, SM cycles} of CTA 0, one record per launch (nullptr = off)
};

constexpr int CONV_STAGES = 4;

template <int CW, int CO_TILE, int TT, int KC>
struct ConvCfg {
    static constexpr int NWC = (CO_TILE / 32) * (TT / 64);   // consumer warps (32 co x 64 t each)
    static constexpr int NTC = NWC * 32;
    static constexpr int NTHREADS = NTC + 32;                // + producer warp
    static constexpr int PITCH = CW + 4;                     // +4 floats: conflict-free float4 row reads
    static constexpr int CHUNK = KC * CO_TILE;               // floats per weight stage
    static constexpr size_t smem_bytes(int window_rows) {
        return 128 + sizeof(float) * ((size_t)CONV_STAGES * CHUNK + (size_t)window_rows * PITCH);
    }
};

template <int CO_TILE, int KC, int PITCH>
__device__ __forceinline__ void mma_chunk(float (&acc)[8][8], const float* __restrict__ xb, const float* __restrict__ wb) {
#pragma unroll 2
    for (int kk = 0; kk < KC; kk += 4) {
        float4 xv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) xv[j] = *reinterpret_cast<const float4*>(xb + j * 8 * PITCH + kk);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 wa = *reinterpret_cast<const float4*>(wb + (kk + q) * CO_TILE);
            const float4 wc = *reinterpret_cast<const float4*>(wb + (kk + q) * CO_TILE + 4);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xq = q == 0 ? xv[j].x : q == 1 ? xv[j].y : q == 2 ? xv[j].z : xv[j].w;
                acc[j][0] = fmaf(xq, wa.x, acc[j][0]);
                acc[j][1] = fmaf(xq, wa.y, acc[j][1]);
                acc[j][2] = fmaf(xq, wa.z, acc[j][2]);
                acc[j][3] = fmaf(xq, wa.w, acc[j][3]);
                acc[j][4] = fmaf(xq, wc.x, acc[j][4]);
                acc[j][5] = fmaf(xq, wc.y, acc[j][5]);
                acc[j][6] = fmaf(xq, wc.z, acc[j][6]);
                acc[j][7] = fmaf(xq, wc.w, acc[j][7]);
            }
        }
    }
}

template <int CW, int CO_TILE, int TT, int KC, bool FUSE>
__global__ void __launch_bounds__(ConvCfg<CW, CO_TILE, TT, KC>::NTHREADS)
conv_gemm_kernel(const ConvArgs a) {
    using Cfg = ConvCfg<CW, CO_TILE, TT, KC>;
    constexpr int NWC = Cfg::NWC, NTC = Cfg::NTC, PITCH = Cfg::PITCH, CHUNK = Cfg::CHUNK;
    constexpr int CO_WARPS = CO_TILE / 32;
    static_assert(CW % KC == 0 && KC % 4 == 0, "bad KC");
    static_assert(!FUSE || CW == CO_TILE, "residual-unit fusion needs Cin == Cout == tile");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + CONV_STAGES;
    float* wst = reinterpret_cast<float*>(smem_raw + 128);
    float* xs = wst + CONV_STAGES * CHUNK;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int j0 = blockIdx.x * TT;                 // first output row of this tile
    const int g = blockIdx.y / a.n_co_tiles;
    const int co_tile = blockIdx.y - g * a.n_co_tiles;
    const int b = blockIdx.z;
    const int n1 = a.n_pieces * a.Ktaps * (CW / KC);     // weight chunks of GEMM 1
    const int ntot = n1 + (FUSE ? CO_TILE / KC : 0);

    if (tid == 0) {
        for (int s = 0; s < CONV_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NWC);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == NWC) {
        // ------------------------------ producer warp: stream weight chunks L2 -> smem (TMA 1-D)
        if (lane == 0) {
            const float* w1 = a.w + (long long)blockIdx.y * a.w_tile_floats;
            for (int c = 0; c < ntot; ++c) {
                const int s = c % CONV_STAGES, it = c / CONV_STAGES;
                if (it > 0) mbar_wait(&empty[s], (it - 1) & 1);
                const float* src = (c < n1) ? w1 + (long long)c * CHUNK : a.w2 + (long long)(c - n1) * CHUNK;
                mbar_arrive_expect_tx(&full[s], CHUNK * 4);
                bulk_g2s(wst + s * CHUNK, src, CHUNK * 4, &full[s]);
            }
        }
        return;
    }

    // ---------------------------------- consumer warps
    const int warp_co = warp % CO_WARPS, warp_t = warp / CO_WARPS;
    const int cg = lane & 3, tg = lane >> 2;
    const int wrows = TT + (a.Ktaps - 1) * a.dil;   // window rows
    float acc[8][8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;

    const float* xg = a.x + (long long)b * a.x_bs + g * a.x_goff;
    const float* sg = a.st_in + (long long)b * a.P * a.st_ld + g * a.st_goff;
    int c = 0;
    for (int piece = 0; piece < a.n_pieces; ++piece) {
        if (piece > 0) named_bar_sync(1, NTC);
        // ---- window load: x~ rows -> smem, history from state, pre-activation applied once
        const int nvec = wrows * (CW / 4);
        for (int idx = tid; idx < nvec; idx += NTC) {
            const int m = idx / (CW / 4);
            const int c4 = idx - m * (CW / 4);
            const int q = piece * CW + c4 * 4;
            int r = 0, ci = q;
            if (a.RG > 1) { r = q >> a.lgCin; ci = q & (a.Cin - 1); }
            const long long i = (long long)(j0 + m) * a.RG + r;     // x~ row
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            long long t = i - a.P;
            if (a.hist_rep && t < 0) t = 0;
            if (t < 0) {
                v = *reinterpret_cast<const float4*>(sg + i * a.st_ld + ci);
            } else {
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
            *reinterpret_cast<float4*>(xs + m * PITCH + c4 * 4) = v;
        }
        named_bar_sync(1, NTC);
        // ---- GEMM 1 over (tap, ci-chunk) of this piece
        for (int tap = 0; tap < a.Ktaps; ++tap) {
            const float* xrow = xs + (warp_t * 64 + tg + tap * a.dil) * PITCH;
            for (int kc0 = 0; kc0 < CW; kc0 += KC, ++c) {
                const int s = c % CONV_STAGES;
                mbar_wait(&full[s], (c / CONV_STAGES) & 1);
                mma_chunk<CO_TILE, KC, PITCH>(acc, xrow + kc0, wst + s * CHUNK + warp_co * 32 + cg * 8);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
            }
        }
    }

    if (FUSE) {
        // ---- residual unit: mid = act(conv_k7(act(x))) stays in smem, then the 1x1 conv
        named_bar_sync(1, NTC);     // everyone is done reading the window
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float* mrow = xs + (warp_t * 64 + tg + 8 * j) * PITCH + warp_co * 32 + cg * 8;
            float4 v0 = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
            float4 v1 = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
            *reinterpret_cast<float4*>(mrow) = apply_act(v0, a.mid_act, a.slope);
            *reinterpret_cast<float4*>(mrow + 4) = apply_act(v1, a.mid_act, a.slope);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
        }
        named_bar_sync(1, NTC);
        const float* xrow = xs + (warp_t * 64 + tg) * PITCH;
        for (int kc0 = 0; kc0 < CO_TILE; kc0 += KC, ++c) {
            const int s = c % CONV_STAGES;
            mbar_wait(&full[s], (c / CONV_STAGES) & 1);
            mma_chunk<CO_TILE, KC, PITCH>(acc, xrow + kc0, wst + s * CHUNK + warp_co * 32 + cg * 8);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
    }

    // ---------------------------------- epilogue: bias, residual, store
    {
        const int co_l = co_tile * CO_TILE + warp_co * 32 + cg * 8;   // channel within the group
        float bv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) bv[i] = a.bias ? a.bias[g * a.Cout_g + co_l + i] : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int t = j0 + warp_t * 64 + tg + 8 * j;
            if (t >= a.Tout) continue;
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = acc[j][i] + bv[i];
            if (a.res) {
                const float* rp = a.res + (long long)b * a.res_bs + (long long)t * a.ldr + g * a.r_goff + co_l;
                const float4 r0 = __ldg(reinterpret_cast<const float4*>(rp));
                const float4 r1 = __ldg(reinterpret_cast<const float4*>(rp + 4));
                // x + y  (residual_unit.py:81: `return x + y`)
                o[0] = r0.x + o[0]; o[1] = r0.y + o[1]; o[2] = r0.z + o[2]; o[3] = r0.w + o[3];
                o[4] = r1.x + o[4]; o[5] = r1.y + o[5]; o[6] = r1.z + o[6]; o[7] = r1.w + o[7];
            }
            if (a.out_nct) {
                float* yp = a.y + (long long)b * a.y_bs + (long long)(g * a.y_goff + co_l) * a.Tout + t;
#pragma unroll
                for (int i = 0; i < 8; ++i) yp[(long long)i * a.Tout] = o[i];
            } else {
                float* yp = a.y + (long long)b * a.y_bs + (long long)t * a.ldy + g * a.y_goff + co_l;
                *reinterpret_cast<float4*>(yp) = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4*>(yp + 4) = make_float4(o[4], o[5], o[6], o[7]);
            }
        }
    }

    // ---------------------------------- new causal state = last P rows of x~ (conv_layer.py:155)
    if (blockIdx.x == gridDim.x - 1 && co_tile == 0 && g < a.st_groups && a.P > 0) {
        float* so = a.st_out + (long long)b * a.P * a.st_ld + g * a.st_goff;
        const int nvec = a.P * (a.Cin / 4);
        for (int idx = tid; idx < nvec; idx += NTC) {
            const int r = idx / (a.Cin / 4);
            const int ci = (idx - r * (a.Cin / 4)) * 4;
            const long long i = (long long)a.T + r;      // x~ row
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

// ------------------------------------------------------------------------------------------------
// stem: Cin = 1 -> COUT, K taps, stride 1 (encoder.py:106-111).  x (B,T) -> y (B,T,COUT).
// ------------------------------------------------------------------------------------------------
struct StemArgs {
    const float* x; long long x_bs;
    const float* st_in; float* st_out;   // (B, K-1)
    int T;
    const float* w;                      // [K][COUT]
    const float* bias;                   // [COUT] or nullptr
    float* y; long long y_bs;
};

template <int COUT, int K>
__global__ void __launch_bounds__(256) stem_kernel(const StemArgs a) {
    constexpr int TT = 1024, P = K - 1, Q = COUT / 4;
    __shared__ float xw[TT + P];
    __shared__ __align__(16) float sw[K * COUT];
    __shared__ __align__(16) float sb[COUT];
    const int b = blockIdx.y, j0 = blockIdx.x * TT, tid = threadIdx.x;
    const float* xg = a.x + (long long)b * a.x_bs;
    for (int i = tid; i < TT + P; i += 256) {
        const long long r = (long long)j0 + i;       // x~ row
        float v = 0.f;
        if (r < P) v = a.st_in[b * P + r];
        else if (r - P < a.T) v = __ldg(xg + r - P);
        xw[i] = v;
    }
    for (int i = tid; i < K * COUT; i += 256) sw[i] = a.w[i];
    for (int i = tid; i < COUT; i += 256) sb[i] = a.bias ? a.bias[i] : 0.f;
    __syncthreads();
    const int q = tid % Q, tl = tid / Q;
    constexpr int TSTEP = 256 / Q;
    float* yg = a.y + (long long)b * a.y_bs;
    for (int t = tl; t < TT; t += TSTEP) {
        if (j0 + t >= a.T) break;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float xv = xw[t + k];
            const float4 w4 = *reinterpret_cast<const float4*>(sw + k * COUT + q * 4);
            acc.x = fmaf(xv, w4.x, acc.x); acc.y = fmaf(xv, w4.y, acc.y);
            acc.z = fmaf(xv, w4.z, acc.z); acc.w = fmaf(xv, w4.w, acc.w);
        }
        const float4 b4 = *reinterpret_cast<const float4*>(sb + q * 4);
        acc.x += b4.x; acc.y += b4.y; acc.z += b4.z; acc.w += b4.w;
        *reinterpret_cast<float4*>(yg + (long long)(j0 + t) * COUT + q * 4) = acc;
    }
    if (blockIdx.x == gridDim.x - 1) {
        for (int r = tid; r < P; r += 256) {
            const long long i = (long long)a.T + r;
            a.st_out[b * P + r] = (i < P) ? a.st_in[b * P + i] : xg[i - P];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// head: CIN -> 1, K taps (decoder.py:133 / HiFiGAN.py:118-123, :294-296).  x (B,T,CIN) -> y (B,T).
// ------------------------------------------------------------------------------------------------
struct HeadArgs {
    const float* x; long long x_bs; int ldx;
    const float* st_in; float* st_out;   // (B, K-1, CIN)
    int T;
    const float* w;                      // [K][CIN]
    float bias; int pre_act; float slope; int post_tanh;
    float* y; long long y_bs;
};

// Eight lanes share one run of R = 8 consecutive outputs: lane c4 owns channels 4*c4..4*c4+3, loads the R + K - 1 window rows of its
// channel quad straight from global (a row is one coalesced 128-byte segment across the eight lanes; no shared memory, every input row
// is read 14/8 times from L1), applies the pre-activation in registers and accumulates its 4-channel share of the R dot products with
// the K x 4 weights it keeps in registers; three xor-shuffles add the eight shares and lane j stores output j.  The layer moves
// 128 B per output row and does 2*K*CIN flops on it: memory-bound once the LDS traffic of the round-1 version (112 LDS.128 per output)
// is gone.
template <int CIN, int K>
__global__ void __launch_bounds__(256) head_kernel(const HeadArgs a) {
    static_assert(CIN == 32, "eight lanes x four channels");
    constexpr int R = 8, P = K - 1, TT = 256;
    const int b = blockIdx.y, tid = threadIdx.x, c4 = tid & 7;
    const int t0 = blockIdx.x * TT + (tid >> 3) * R;               // first output of this lane group
    const float* xg = a.x + (long long)b * a.x_bs;
    const float* sg = a.st_in + (long long)b * P * CIN;
    float4 w[K];
#pragma unroll
    for (int k = 0; k < K; ++k) w[k] = __ldg(reinterpret_cast<const float4*>(a.w + k * CIN) + c4);
    float acc[R];
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = 0.f;
    if (t0 < a.T) {
        float4 xv[R + P];
#pragma unroll
        for (int r = 0; r < R + P; ++r) {
            const long long i = (long long)t0 + r;                  // x~ row = history(P) || chunk
            xv[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < P) xv[r] = __ldg(reinterpret_cast<const float4*>(sg + i * CIN) + c4);            // state rows are stored post-activation
            else if (i - P < a.T) xv[r] = apply_act(__ldg(reinterpret_cast<const float4*>(xg + (i - P) * a.ldx) + c4), a.pre_act, a.slope);
        }
#pragma unroll
        for (int j = 0; j < R; ++j)
#pragma unroll
            for (int k = 0; k < K; ++k) {
                acc[j] = fmaf(xv[j + k].x, w[k].x, acc[j]); acc[j] = fmaf(xv[j + k].y, w[k].y, acc[j]);
                acc[j] = fmaf(xv[j + k].z, w[k].z, acc[j]); acc[j] = fmaf(xv[j + k].w, w[k].w, acc[j]);
            }
    }
    float mine = 0.f;
#pragma unroll
    for (int j = 0; j < R; ++j) {
        float v = acc[j];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        if (j == c4) mine = v;
    }
    if (t0 + c4 < a.T) {
        mine += a.bias;
        if (a.post_tanh) mine = tanhf(mine);
        a.y[(long long)b * a.y_bs + t0 + c4] = mine;
    }
    if (blockIdx.x == gridDim.x - 1) {
        float* so = a.st_out + (long long)b * P * CIN;
        for (int idx = tid; idx < P * (CIN / 4); idx += 256) {
            const int r = idx / (CIN / 4), ci = (idx - r * (CIN / 4)) * 4;
            const long long i = (long long)a.T + r;
            float4 v;
            if (i < P) v = *reinterpret_cast<const float4*>(sg + i * CIN + ci);
            else v = apply_act(__ldg(reinterpret_cast<const float4*>(xg + (i - P) * a.ldx + ci)), a.pre_act, a.slope);
            *reinterpret_cast<float4*>(so + r * CIN + ci) = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// residual VQ (layers/vq_module.py:90-104,136-149).  Rounding order == oracle/rvq_oracle.c.
// ------------------------------------------------------------------------------------------------
struct RvqArgs {
    const float* z;        // (B, D, F) channels-first (what encode() returns)
    int B, F, nq;
    const float* embed;    // (nq, D, N) = state-dict `embed` tensors stacked
    const float* e2;       // (nq, N)   ||e||^2 in torch's summation order
    long long* idx;        // (nq, B, F) flat indices (+ N*i), or nullptr
    // fused outputs (bin/stream.py:224 hand-off): either may be nullptr
    unsigned char* packed; // (B*F, bpf) index bitstream, same format as pack_kernel
    float* zq;             // (B*F, D) = lookup(idx) (vq_module.py:159-161), so that quantize can hand zq straight to the decoder
    int bits, bpf, n_pass; // bits per index, bytes per packed frame, frame passes per block (frames per block = n_pass * FP)
};

constexpr int RVQ_THREADS = 256;

// One block quantises n_pass * FP frames through all nq stages.  Thread t owns codewords t, t + 256, ... (NPT of them) for FP frames
// at a time: FP * NPT independent fused-multiply-add chains over k (ascending, one chain per distance: MKL's sgemm order, which is
// what makes the indices bit-identical to torch-CPU), codeword elements streamed from L2 (the 256 KB stage table is shared by all
// blocks), residuals broadcast from shared memory.  The host picks (FP, n_pass) so that the grid is one wave with the least idle
// tail (adec_quantize), which alone took the 64 x 160-frame case from 0.58 ms (320 blocks of 32 frames = 1.08 waves) to one wave.
template <int D, int NPT, int FP>   // codebook size N = RVQ_THREADS * NPT
__global__ void __launch_bounds__(RVQ_THREADS) rvq_kernel(const RvqArgs a) {
    constexpr int N = RVQ_THREADS * NPT, NW = RVQ_THREADS / 32;
    static_assert(D % 32 == 0 && FP % 4 == 0, "D must be a multiple of 32, FP of 4");
    extern __shared__ __align__(16) float rvq_smem[];
    const int FR = a.n_pass * FP;
    float* r = rvq_smem;                               // [FR][D] residuals
    float* zq = r + FR * D;                            // [FR][D] sum of the chosen codewords
    float* r2 = zq + FR * D;                           // [FR][D] 2 * residual (exact): the multiplicand of the distance GEMM, kept so that the
                                                       //         inner loop is FFMA + one broadcast LDS.128 per 16 of them (was + 4 FMUL)
    float* x2 = r2 + FR * D;                           // [FR]
    float* wv = x2 + FR;                               // [FR][NW]
    int* wi = reinterpret_cast<int*>(wv + FR * NW);    // [FR][NW]
    int* best = wi + FR * NW;                          // [nq][FR] local indices of every stage (for the packed frame)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long nfr = (long long)a.B * a.F;
    const long long f0 = (long long)blockIdx.x * FR;
    for (int i = tid; i < FR * D; i += RVQ_THREADS) {
        const int f = i % FR, k = i / FR;            // consecutive threads -> consecutive frames: coalesced reads of z (B,D,F)
        const long long fr = f0 + f;
        float v = 0.f;
        if (fr < nfr) {
            const long long bb = fr / a.F, ff = fr - bb * a.F;
            v = a.z[(bb * D + k) * a.F + ff];        // quantizer.py:43 z.transpose(2,1)
        }
        r[f * D + k] = v;
        r2[f * D + k] = 2.0f * v;
        zq[f * D + k] = 0.f;
    }
    __syncthreads();
    for (int st = 0; st < a.nq; ++st) {
        const float* E = a.embed + (long long)st * D * N;
        // x2 = flatten.pow(2).sum(1): 8-lane vectors, 4 interleaved accumulators, sequential horizontal add
        for (int f = tid; f < FR; f += RVQ_THREADS) {
            float accv[4][8];
#pragma unroll
            for (int v = 0; v < 4; ++v)
#pragma unroll
                for (int l = 0; l < 8; ++l) accv[v][l] = 0.f;
#pragma unroll
            for (int v = 0; v < D / 8; ++v)
#pragma unroll
                for (int l = 0; l < 8; ++l) {
                    const float xv = r[f * D + 8 * v + l];
                    accv[v & 3][l] = __fadd_rn(accv[v & 3][l], __fmul_rn(xv, xv));
                }
            float s = 0.f;
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                const float t = __fadd_rn(__fadd_rn(__fadd_rn(accv[0][l], accv[1][l]), accv[2][l]), accv[3][l]);
                s = (l == 0) ? t : __fadd_rn(s, t);
            }
            x2[f] = s;
        }
        float e2v[NPT];
#pragma unroll
        for (int m = 0; m < NPT; ++m) e2v[m] = __ldg(a.e2 + (long long)st * N + tid + m * RVQ_THREADS);
        __syncthreads();   // x2 visible
        // dot2[c] = sum_k (2 r_k) * E[k][c], k ascending, one fused multiply-add chain per output (MKL sgemm order)
#pragma unroll 1
        for (int fh = 0; fh < FR; fh += FP) {
            float acc[FP][NPT];
#pragma unroll
            for (int f = 0; f < FP; ++f)
#pragma unroll
                for (int m = 0; m < NPT; ++m) acc[f][m] = 0.f;
            float e[4][NPT], en[4][NPT];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int m = 0; m < NPT; ++m) e[kk][m] = __ldg(E + (long long)kk * N + tid + m * RVQ_THREADS);
#pragma unroll 1
            for (int k = 0; k < D; k += 4) {
                const int kn = k + 4 < D ? k + 4 : k;            // prefetch the next four codeword rows while these are used
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int m = 0; m < NPT; ++m) en[kk][m] = __ldg(E + (long long)(kn + kk) * N + tid + m * RVQ_THREADS);
#pragma unroll
                for (int f = 0; f < FP; ++f) {
                    const float4 r4 = *reinterpret_cast<const float4*>(&r2[(fh + f) * D + k]);
                    const float rk[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                        for (int m = 0; m < NPT; ++m) acc[f][m] = fmaf(rk[kk], e[kk][m], acc[f][m]);
                }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int m = 0; m < NPT; ++m) e[kk][m] = en[kk][m];
            }
#pragma unroll
            for (int f = 0; f < FP; ++f) {
                // dist = (x2 - dot2) + e2 ; index = first arg-max of -dist  (vq_module.py:93-98)
                float bv = 0.f;
                int bi = 0;
#pragma unroll
                for (int m = 0; m < NPT; ++m) {
                    const float nd = -__fadd_rn(__fsub_rn(x2[fh + f], acc[f][m]), e2v[m]);
                    if (m == 0 || nd > bv) { bv = nd; bi = tid + m * RVQ_THREADS; }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                if (lane == 0) { wv[(fh + f) * NW + warp] = bv; wi[(fh + f) * NW + warp] = bi; }
            }
        }
        __syncthreads();
        for (int f = tid; f < FR; f += RVQ_THREADS) {
            float bv = wv[f * NW];
            int bi = wi[f * NW];
#pragma unroll
            for (int w = 1; w < NW; ++w) {
                const float ov = wv[f * NW + w];
                const int oi = wi[f * NW + w];
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            best[st * FR + f] = bi;
            const long long fr = f0 + f;
            if (a.idx && fr < nfr) a.idx[(long long)st * nfr + fr] = (long long)bi + (long long)N * st;   // vq_module.py:145-146
        }
        __syncthreads();
        // quantize = x + (e - x); residual -= quantize  (vq_module.py:101-102,143); zq += e in stage order (vq_module.py:159-161)
        for (int i = tid; i < FR * D; i += RVQ_THREADS) {
            const int f = i / D, k = i - f * D;
            const float rv = r[i];
            const float q = __ldg(E + (long long)k * N + best[st * FR + f]);
            const float qq = __fadd_rn(rv, __fsub_rn(q, rv));
            const float rn = __fsub_rn(rv, qq);
            r[i] = rn;
            r2[i] = 2.0f * rn;
            zq[i] = st == 0 ? q : __fadd_rn(zq[i], q);
        }
        __syncthreads();
    }
    if (a.zq) {
        for (int i = tid; i < FR * D / 4; i += RVQ_THREADS) {
            const long long fr = f0 + (i * 4) / D;
            if (fr < nfr) *reinterpret_cast<float4*>(a.zq + f0 * D + (long long)i * 4) = *reinterpret_cast<const float4*>(zq + i * 4);
        }
    }
    if (a.packed) {
        // packed frame: nq local indices of `bits` bits each, stage 0 first, little-endian bit order (same bytes as pack_kernel)
        for (int f = tid; f < FR; f += RVQ_THREADS) {
            const long long fr = f0 + f;
            if (fr >= nfr) continue;
            unsigned char* o = a.packed + fr * a.bpf;
            unsigned long long accb = 0;
            int nb = 0, ob = 0;
            for (int i = 0; i < a.nq; ++i) {
                accb |= (unsigned long long)best[i * FR + f] << nb;
                nb += a.bits;
                while (nb >= 8) { o[ob++] = (unsigned char)(accb & 0xffu); accb >>= 8; nb -= 8; }
            }
            if (nb > 0) o[ob++] = (unsigned char)(accb & 0xffu);
        }
    }
}

// codebook lookup (vq_module.py:159-161): zq[b][f][:] = sum_i codebook[idx[i][b][f]][:], i ascending.  The indices come either as the
// int64 (nq, B*F) tensor of the reference or straight from the packed bitstream (unpack fused into the lookup).
struct LookupArgs {
    const long long* idx;   // (nq, B*F) or nullptr
    const unsigned char* packed;   // (B*F, bpf) or nullptr
    long long nfr;
    int nq, D, N, bits, bpf;
    const float* codebook;  // (nq*N, D)
    long long n_rows;       // nq*N (bounds check)
    float* zq;              // (B*F, D)
    int* err;               // bit 0 set on an out-of-range index
};

__global__ void __launch_bounds__(256) lookup_kernel(const LookupArgs a) {
    const int vpf = a.D / 4;    // float4 per frame
    const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
    if (gid >= a.nfr * vpf) return;
    const long long fr = gid / vpf;
    const int k4 = (int)(gid - fr * vpf) * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    const unsigned char* in = a.packed ? a.packed + fr * a.bpf : nullptr;
    const unsigned long long mask = (1ull << a.bits) - 1ull;
    unsigned long long accb = 0;
    int nb = 0, ib = 0;
    for (int i = 0; i < a.nq; ++i) {
        long long row;
        if (in) {
            while (nb < a.bits) { accb |= (unsigned long long)in[ib++] << nb; nb += 8; }
            const long long v = (long long)(accb & mask);
            accb >>= a.bits; nb -= a.bits;
            row = v < a.N ? v + (long long)i * a.N : -1;
        } else {
            row = a.idx[(long long)i * a.nfr + fr];
        }
        if (row < 0 || row >= a.n_rows) { atomicOr(a.err, 1); continue; }
        const float4 v = __ldg(reinterpret_cast<const float4*>(a.codebook + row * a.D + k4));
        if (i == 0) s = v;
        else { s.x = __fadd_rn(s.x, v.x); s.y = __fadd_rn(s.y, v.y); s.z = __fadd_rn(s.z, v.z); s.w = __fadd_rn(s.w, v.w); }
    }
    *reinterpret_cast<float4*>(a.zq + fr * a.D + k4) = s;
}

// Index bitstream.  The reference has no wire format: it ships int64 (Nq,F) tensors through a queue
// (bin/stream.py:224).  Packed frame = Nq local indices (idx - i*N) of `bits` = ceil(log2 N) bits each, stage 0 first, little-endian
// bit order, zero-padded to whole bytes: 8 x 10 bit = 10 bytes per frame.  One thread per frame (a frame is 10-20 bytes).
struct PackArgs {
    long long* idx;           // (nq, nfr) flat indices (+N*i), read by pack / written by unpack
    unsigned char* packed;    // (nfr, bpf)
    long long nfr;
    int nq, N, bits, bpf;
    int* err;                 // set to 1 on an out-of-range index / code
};

__global__ void __launch_bounds__(256) pack_kernel(const PackArgs a) {
    const long long fr = (long long)blockIdx.x * 256 + threadIdx.x;
    if (fr >= a.nfr) return;
    unsigned char* o = a.packed + fr * a.bpf;
    unsigned long long acc = 0;
    int nb = 0, ob = 0;
    for (int i = 0; i < a.nq; ++i) {
        long long v = a.idx[(long long)i * a.nfr + fr] - (long long)i * a.N;
        if (v < 0 || v >= a.N) { atomicOr(a.err, 1); v = 0; }
        acc |= (unsigned long long)v << nb;
        nb += a.bits;
        while (nb >= 8) { o[ob++] = (unsigned char)(acc & 0xffu); acc >>= 8; nb -= 8; }
    }
    if (nb > 0) o[ob++] = (unsigned char)(acc & 0xffu);
}

__global__ void __launch_bounds__(256) unpack_kernel(const PackArgs a) {
    const long long fr = (long long)blockIdx.x * 256 + threadIdx.x;
    if (fr >= a.nfr) return;
    const unsigned char* in = a.packed + fr * a.bpf;
    const unsigned long long mask = (1ull << a.bits) - 1ull;
    unsigned long long acc = 0;
    int nb = 0, ib = 0;
    for (int i = 0; i < a.nq; ++i) {
        while (nb < a.bits) { acc |= (unsigned long long)in[ib++] << nb; nb += 8; }
        long long v = (long long)(acc & mask);
        acc >>= a.bits; nb -= a.bits;
        if (v >= a.N) { atomicOr(a.err, 1); v = 0; }
        a.idx[(long long)i * a.nfr + fr] = v + (long long)i * a.N;
    }
}

// replicate stream 0's state to all streams (adec_set_streams)
__global__ void replicate_kernel(float* dst, const float* src, long long per_stream, int n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < per_stream * n) dst[i] = src[i % per_stream];
}

}  // namespace adec
