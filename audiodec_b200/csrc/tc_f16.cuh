// tc_conv_f16_kernel: the causal conv family on tcgen05 `kind::f16` tensor cores, persistent, sm_100a (round 2 default engine).
//
// Same ConvArgs contract and the same schedule as the round-1 3xTF32 kernel (tc_persist.cuh): GEMM orientation M = 128 time
// steps (TMEM lanes), N = output-channel tile NT (TMEM columns), K = input channels of one tap; both operands K-major,
// no-swizzle "column blocks"  smem[(kb * ROWS + row) * 16 B]  (kb = 16-byte block of 8 channels, SBO = 128 B, LBO = ROWS*16 B),
// so a conv tap is a row-shifted start address of the same window - no im2col, no copies.  What changed is the arithmetic:
//
//   PREC = 3 (fp32-grade, every layer that must stay within 1e-4 / bit-identical indices)
//       a     = A_hi + 2^-11 A_lo         A_hi = fp16(a),      A_lo = fp16((a - A_hi) * 2^11)        (producer warps)
//       w * s = W_hi + W_lo               W_hi = fp16(w * s),  W_lo = fp16(w * s - W_hi),  W_his = 2^-11 W_hi   (host, s = 2^p per op:
//                                                                                            max|w| s in [2^12, 2^13) keeps all three normal)
//       a w s ~= A_lo W_his + A_hi W_lo + A_hi W_hi     three fp16 products, fp32 accumulation in TMEM, result * 2^-p in the drain warps.
//     fp16 and tf32 both carry 11 significand bits, so this is as accurate as 3xTF32 (dropped term ~2^-22 relative; measured in
//     tools/tc_probe3.cu: max error 6.5e-7 of the output rms at K = 224, below a plain fp32 FMA chain's 1.8e-6) - but one fp16 MMA
//     covers K = 16 where a tf32 MMA covers K = 8 for the same issue slot and the same shared-memory operand bytes: half the
//     tensor-pipe time and half the operand reads per conv.  Range: |a| < 65504 (checked in the epilogue, ConvArgs::err).
//   PREC = 1 (bf16 operands, one product; the HiFi-GAN vocoder's bf16 mode, BASELINE configs[2])
//
// The TMEM accumulator still truncates at every accumulation step (tools/tc_probe2.cu, tc_probe3.cu: -2e-7 of the rms over 14 chained
// steps), so accumulation stays GROUPED: one group = one 32-channel piece x TWO taps = 4 full-magnitude MMAs (preceded by the 8
// small-term MMAs) into a fresh TMEM partial; drain warps add the partials into fp32 registers with round-to-nearest adds
// (bias -6.7e-8 of the rms).  Groups are twice as long as in round 1 (K = 64), so the drain work per conv halves too.
//
// Warp roles, barriers and the barrier rule are those of tc_persist.cuh: warp 0 = weight producer (1-D TMA bulk
// copies of host-packed stages) + TMEM allocator, warps 1..NW = MMA issuers taking groups round robin, warps 4.. = activation
// producers (global -> pre-activation -> hi/lo fp16 split -> smem), last 8 warps = drain + fused intermediate + epilogue.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "tc_kernels.cuh"

namespace adec {

// -DADEC_TIMELINE: CTA 1 records {event << 24 | index, clock} pairs into ConvArgs::tl (tools/timeline_f16.py prints them)
#ifdef ADEC_TIMELINE
// every recording thread owns a region of 8192 records and a private counter: plain stores, no atomics (an atomic's round trip would
// stall the recording warp for ~600 cycles per event)
#define ADEC_TL_DECL(role) unsigned int tl_n = 0; unsigned int* const tl_base = a.tl ? a.tl + 2 + 2 * 8192 * (role) : nullptr
#define ADEC_TL(code, idx)                                                                        \
    do {                                                                                          \
        if (tl_base && blockIdx.x == 1 && tl_n < 8192u) {                                         \
            tl_base[2 * tl_n] = ((unsigned)(code) << 24) | ((unsigned)(idx) & 0xffffffu); tl_base[2 * tl_n + 1] = (unsigned)clock64(); ++tl_n; \
        }                                                                                         \
    } while (0)
#else
#define ADEC_TL_DECL(role) do { } while (0)
#define ADEC_TL(code, idx) do { } while (0)
#endif

constexpr int F16_KB = TC_CP / 8;      // 16-byte K blocks (8 channels) per 32-channel piece and plane
constexpr int F16_MIDP = 128;          // rows of the fused intermediate operand (one row per drain lane: 16-byte stores are conflict-free)
constexpr float F16_LO_SCALE = 2048.f; // 2^11

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Same MMA with the two shared-memory descriptors given by their LOW words (start address >> 4 | LBO >> 4 << 16); the high word
// (SBO = 128 B, descriptor version 1) is the constant 0x4008.  The issuing lane advances the low words by small additions, which keeps
// its instruction count per MMA at ~4 uniform-datapath instructions (umma_desc() from scratch: ~11 - at N <= 64 the issue rate of
// one lane, ~80 cycles per MMA, was below what the tensor pipe accepts, 43 / 51 cycles: tools/probe_align.py).
__device__ __forceinline__ void umma_f16_lo(uint32_t tmem_d, uint32_t a_lo_word, uint32_t b_lo_word, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n .reg .b64 da, db;\n setp.ne.b32 p, %4, 0;\n mov.b64 da, {%1, %5};\n mov.b64 db, {%2, %5};\n"
        " tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}" ::"r"(tmem_d),
        "r"(a_lo_word), "r"(b_lo_word), "r"(idesc), "r"(accumulate), "r"(0x4008u)
        : "memory");
}
__device__ __forceinline__ uint32_t umma_lo_word(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr & 0x3FFFF) >> 4) | (((lbo_bytes >> 4) & 0x3FFF) << 16); }

template <int NT, int PREC> struct TcfCfg {
    // One CTA per SM.  (Tried at NT = 32: two CTAs of 4 producer + 4 drain warps per SM - 0.38 -> 0.52 ms on the C = 32 unit, the
    // 80-register budget spills and each pipeline has half the warps.)
    static constexpr int NPR = PREC == 3 ? 3 : 1;                  // weight planes per tap: hi | lo | hi * 2^-11
    static constexpr int NPL = PREC == 3 ? 2 : 1;                  // activation planes: hi | lo * 2^11
    static constexpr int TAP_BYTES = NPR * F16_KB * NT * 16;       // one (piece, tap) of weights
    static constexpr int STAGE_BYTES = 2 * TAP_BYTES;              // one group = up to two taps
    static constexpr int STAGES = NT == 64 ? 4 : 3;
    static constexpr int NW = (STAGES % 2) ? 3 : 2;                // MMA issuer warps (NW | STAGES, NW | NPB: the barrier rule)
    static constexpr int NPB = NT == 128 ? (NW == 3 ? 3 : 4) : (NW == 3 ? 6 : 8);   // TMEM partials
    static constexpr int MB = NT == 64 ? 2 : 1;                    // fused-intermediate buffers
    static constexpr int NDG = 2;                                  // drain groups (NT=32: each owns half a 32-column piece)
    // Producer threads: 8 warps wherever the 96-register budget of a 640-thread CTA holds the drain warps' accumulators - the producers
    // are bound by the global loads they can keep in flight (ncu of the first f16 build, NT = 128 with 4 producer warps: long-scoreboard
    // 10.5 stalls per issue, issue slots 22 % busy).  The fused NT = 128 unit (64 + 64 live accumulators) keeps 4 producer warps and
    // 128 registers.  (setmaxnreg re-balancing was tried: ptxas caps the control / producer sections as asked but does not give the
    // drain section more than the launch bound, so it only added spills.)
#ifndef ADEC_FUSE_TEAMS
#define ADEC_FUSE_TEAMS 0        // 1 = the fused NT = 32 / 64 units also run two producer teams (A/B)
#endif
#ifndef ADEC_UNR_E
#define ADEC_UNR_E 2
#endif
#ifndef ADEC_NT128_PLAIN_NPROD
#define ADEC_NT128_PLAIN_NPROD 256
#endif
    __host__ __device__ static constexpr int nprod(bool fuse) { return NT == 128 ? (fuse ? 128 : ADEC_NT128_PLAIN_NPROD) : 256; }
    __host__ __device__ static constexpr int threads(bool fuse) { return 128 + nprod(fuse) + 128 * NDG; }
    static constexpr int MID_BYTES = NPL * F16_KB * F16_MIDP * 16; // one intermediate piece (32 channels)
    __host__ __device__ static constexpr int win_pitch(int wrows) { return ((wrows + 1) & ~3) + 2; }      // rows, == 2 mod 4: conflict-free producer stores
    __host__ __device__ static constexpr int win_bytes(int wrows) { return NPL * F16_KB * win_pitch(wrows) * 16; }
    // window buffers: as many as fit (2..4): the producers run that many pieces ahead of the MMAs
    static int n_wbuf(int wrows, bool fuse) {
        const long long avail = 227 * 1024 - 512 - (long long)STAGES * STAGE_BYTES - (fuse ? (long long)MB * MID_BYTES : 0);
        const long long n = avail / win_bytes(wrows);
        return (int)(n > 4 ? 4 : n);
    }
    static size_t smem_bytes(int wrows, bool fuse) {
        return 512 + (size_t)STAGES * STAGE_BYTES + (size_t)n_wbuf(wrows, fuse) * win_bytes(wrows) + (fuse ? (size_t)MB * MID_BYTES : 0);
    }
};

// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256): one instruction and one full 32-byte sector per lane where two 128-bit
// accesses would each touch half a sector - the epilogue's row-per-lane stores and skip loads cost one LSU line transaction per lane
// and instruction, so halving the instructions halves that cost.  Addresses must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const float* p, float4& u, float4& v) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(u.x), "=f"(u.y), "=f"(u.z), "=f"(u.w), "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]),
                 "f"(v[6]), "f"(v[7])
                 : "memory");
}

// 8 consecutive channels of one row -> one 16-byte hi block (+ one lo block)
template <int PREC>
__device__ __forceinline__ void split_store(unsigned char* hi, unsigned char* lo, const float4 u, const float4 v) {
    if (PREC == 3) {
        const __half2 h0 = __floats2half2_rn(u.x, u.y), h1 = __floats2half2_rn(u.z, u.w), h2 = __floats2half2_rn(v.x, v.y), h3 = __floats2half2_rn(v.z, v.w);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
        const __half2 l0 = __floats2half2_rn((u.x - f0.x) * F16_LO_SCALE, (u.y - f0.y) * F16_LO_SCALE);
        const __half2 l1 = __floats2half2_rn((u.z - f1.x) * F16_LO_SCALE, (u.w - f1.y) * F16_LO_SCALE);
        const __half2 l2 = __floats2half2_rn((v.x - f2.x) * F16_LO_SCALE, (v.y - f2.y) * F16_LO_SCALE);
        const __half2 l3 = __floats2half2_rn((v.z - f3.x) * F16_LO_SCALE, (v.w - f3.y) * F16_LO_SCALE);
        uint4 H, L;
        H.x = *reinterpret_cast<const uint32_t*>(&h0); H.y = *reinterpret_cast<const uint32_t*>(&h1);
        H.z = *reinterpret_cast<const uint32_t*>(&h2); H.w = *reinterpret_cast<const uint32_t*>(&h3);
        L.x = *reinterpret_cast<const uint32_t*>(&l0); L.y = *reinterpret_cast<const uint32_t*>(&l1);
        L.z = *reinterpret_cast<const uint32_t*>(&l2); L.w = *reinterpret_cast<const uint32_t*>(&l3);
        *reinterpret_cast<uint4*>(hi) = H;
        *reinterpret_cast<uint4*>(lo) = L;
    } else {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(u.x, u.y), h1 = __floats2bfloat162_rn(u.z, u.w), h2 = __floats2bfloat162_rn(v.x, v.y),
                             h3 = __floats2bfloat162_rn(v.z, v.w);
        uint4 H;
        H.x = *reinterpret_cast<const uint32_t*>(&h0); H.y = *reinterpret_cast<const uint32_t*>(&h1);
        H.z = *reinterpret_cast<const uint32_t*>(&h2); H.w = *reinterpret_cast<const uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(hi) = H;
    }
}

// Persistent-tile iterator: tile = xt + n_x * (y + n_y * b) advances by gridDim.x per step.  Decoding it with / and % costs ~100
// instructions per tile in EVERY thread (12 % of all instructions of the C = 32 unit in the first ncu capture); the increments below
// are carried additions.
struct TileIter {
    int tile, xt, y, b, sx, sy, sb, nx, ny;
    __device__ __forceinline__ TileIter(int first, int step, int n_x, int n_y) {
        nx = n_x; ny = n_y; tile = first;
        xt = first % n_x;
        const int q = first / n_x;
        y = q % n_y; b = q / n_y;
        sx = step % n_x;
        const int sq = step / n_x;
        sy = sq % n_y; sb = sq / n_y;
    }
    __device__ __forceinline__ void next(int step) {
        tile += step;
        xt += sx;
        const int cx = xt >= nx ? 1 : 0;
        xt -= cx ? nx : 0;
        y += sy + cx;
        const int cy = y >= ny ? 1 : 0;
        y -= cy ? ny : 0;
        b += sb + cy;
    }
};

__device__ __forceinline__ float4 norm4(float4 x, const float* mean, const float* scale) {
    const float4 mu = *reinterpret_cast<const float4*>(mean);
    const float4 sc = *reinterpret_cast<const float4*>(scale);
    x.x = __fdiv_rn(x.x - mu.x, sc.x); x.y = __fdiv_rn(x.y - mu.y, sc.y);
    x.z = __fdiv_rn(x.z - mu.z, sc.z); x.w = __fdiv_rn(x.w - mu.w, sc.w);
    return x;
}

template <int NT, bool FUSE, int PRE, int PREC>
__global__ void __launch_bounds__(TcfCfg<NT, PREC>::threads(FUSE), 1) tc_conv_f16_kernel(const ConvArgs a, int n_xtiles, int n_ytiles, int n_tiles) {
    using Cfg = TcfCfg<NT, PREC>;
    constexpr int S = Cfg::STAGES, CP = TC_CP, TT = TC_TT, NDG = Cfg::NDG, KB = F16_KB, MIDP = F16_MIDP;
    constexpr int NPROD = Cfg::nprod(FUSE), DRAIN0 = (128 + NPROD) / 32, NPB = Cfg::NPB, MB = Cfg::MB, NW = Cfg::NW;
    constexpr int STAGE_BYTES = Cfg::STAGE_BYTES, TAP_BYTES = Cfg::TAP_BYTES, PLANE_B = KB * NT * 16;   // one weight plane of one tap
    constexpr int NCOL = NT / NDG;                       // accumulator registers per drain thread
    constexpr bool HALF = NCOL < CP;                     // NT=32: a drain group owns 16 of the piece's 32 columns
    constexpr int PPG = HALF ? 1 : NCOL / CP;            // 32-column pieces (or half pieces) owned by one drain group
    constexpr int UC = HALF ? NCOL : CP;                 // columns per owned unit
    constexpr bool PIPE = FUSE && NT == 32;              // conv of tile i+1 issued before the 1x1 conv of tile i
    // instruction descriptor: D = f32 (1 << 4), A/B = f16 (0) or bf16 (1) at bits 7 / 10, both K-major, N >> 3 at 17, M >> 4 at 24
    constexpr uint32_t FMT = PREC == 3 ? 0u : 1u;
    constexpr uint32_t IDESC = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    static_assert(S % NW == 0 && NPB % NW == 0, "a weight stage / TMEM partial must belong to one MMA warp");
    constexpr uint32_t TMEM_COLS = NPB * NT <= 32 ? 32 : NPB * NT <= 64 ? 64 : NPB * NT <= 128 ? 128 : NPB * NT <= 256 ? 256 : 512;
    static_assert(NPB * NT <= 512, "TMEM has 512 columns");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* b_full = reinterpret_cast<uint64_t*>(smem_raw);     // [S]   weights landed
    uint64_t* b_empty = b_full + S;                                // [S]   weights consumed
    uint64_t* w_full = b_empty + S;                                // [4]   window piece written (a.n_wbuf <= 4 buffers in use)
    uint64_t* w_empty = w_full + 4;                                // [4]   window piece consumed
    uint64_t* m_full = w_empty + 4;                                // [MB]  fused-intermediate piece written
    uint64_t* m_empty = m_full + 2;                                // [2]   ... consumed; indexed by the drain group that writes the freed buffer NEXT
    uint64_t* p_full = m_empty + 2;                                // [NPB] TMEM partial complete
    uint64_t* p_empty = p_full + NPB;                              // [NPB] TMEM partial drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_empty + NPB);
    unsigned char* bst = smem_raw + 512;                           // up to 40 barriers + the TMEM slot live in the first 512 B
    const int wrows = TT + (a.Ktaps - 1) * a.dil;
    const int wrp = Cfg::win_pitch(wrows);
    const int win_b = Cfg::win_bytes(wrows);
    unsigned char* wbuf0 = bst + S * STAGE_BYTES;                  // a.n_wbuf window buffers of win_b bytes
    unsigned char* mbuf = wbuf0 + (size_t)a.n_wbuf * win_b;        // FUSE only: MB x MID_BYTES

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    unsigned long long kt0 = 0;
    long long kc0 = 0;
    if (a.dbg && blockIdx.x == 0 && tid == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(kt0)); kc0 = clock64(); }
    const int gpp = (a.Ktaps + 1) >> 1;                            // groups per piece: tap pairs (+ one single tap)
    // TMEM partials per tile.  gspan = 0: one partial per group (4 main accumulation steps, then round-to-nearest register adds);
    // gspan = 1: one partial per 32-channel PIECE (all its taps: 14 main steps at K = 7) and one for the whole 1x1 conv - 3.5x fewer
    // TMEM -> register round trips, which bound the wide layers (timeline: ~2400 cycles of tcgen05.ld latency per partial at NT = 128
    // against 816 cycles of MMAs per group); a partial's MMAs all come from one issuer warp (partials round robin over the warps).
    const bool span = a.gspan != 0;
    const int n_g1 = span ? a.n_pieces : a.n_pieces * gpp;
    const int n_g2 = FUSE ? (span ? 1 : NT / CP) : 0;
    const int n_s2 = FUSE ? NT / CP : 0;                          // weight stages of the 1x1 conv

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], a.gspan ? NW : 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&w_full[i], ((!FUSE || ADEC_FUSE_TEAMS) && NPROD == 256) ? NPROD / a.teams : NPROD); mbar_init(&w_empty[i], NW); }
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
    // programmatic dependent launch (adec.cu launch_tcf): everything above - and the weight stream below, weights being constants - may
    // overlap the previous kernel's tail; the warps that read activations / state / skip tensors or write outputs wait for that grid first
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp >= 4) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (warp == 0) {
        // ------------------------------------------------ weight producer: one bulk copy per group (1 or 2 taps)
        if (lane == 0) {
            int c = 0;
            ADEC_TL_DECL(0);
            auto stream = [&](const unsigned char* base, int pieces, int taps) {
                for (int p = 0; p < pieces; ++p)
                    for (int t0 = 0; t0 < taps; t0 += 2, ++c) {
                        const int s = c % S, it = c / S;
                        const uint32_t bytes = (uint32_t)(taps - t0 >= 2 ? 2 : 1) * TAP_BYTES;
                        if (it > 0) mbar_wait(&b_empty[s], (it - 1) & 1, 100);
                        ADEC_TL(1, c);
                        if (a.dbg_wdiv < 0) { mbar_arrive(&b_full[s]); continue; }
                        const uint32_t lb = a.dbg_wdiv > 1 ? (bytes / (uint32_t)a.dbg_wdiv) & ~15u : bytes;
                        mbar_arrive_expect_tx(&b_full[s], lb);
                        bulk_g2s(bst + s * STAGE_BYTES, base + ((long long)p * taps + t0) * TAP_BYTES, lb, &b_full[s]);
                    }
            };
            const unsigned char* w1 = reinterpret_cast<const unsigned char*>(a.w);
            const unsigned char* w2 = reinterpret_cast<const unsigned char*>(a.w2);
            auto w1_of = [&](int y) { return w1 + (long long)y * a.w_tile_floats * 4; };
            TileIter it(blockIdx.x, gridDim.x, n_xtiles, n_ytiles);
            if (PIPE) {
                // same order as the MMA warps: G1(t0), then per tile { G1(next), G2(this) }
                if (it.tile < n_tiles) stream(w1_of(it.y), a.n_pieces, a.Ktaps);
                while (it.tile < n_tiles) {
                    it.next(gridDim.x);
                    if (it.tile < n_tiles) stream(w1_of(it.y), a.n_pieces, a.Ktaps);
                    stream(w2, n_s2, 1);
                }
            } else {
                for (; it.tile < n_tiles; it.next(gridDim.x)) {
                    stream(w1_of(it.y), a.n_pieces, a.Ktaps);
                    if (FUSE) stream(w2, n_s2, 1);
                }
            }
        }
    } else if (warp >= 1 && warp <= NW) {
        // ------------------------------------------------ MMA issuers (groups round robin)
        const int mw = warp - 1;
        int c = 0, mp = 0;
        ADEC_TL_DECL(1 + mw);
        int wb = 0, wround = 0;                                   // window piece counter wp = wround * n_wbuf + wb
        const uint32_t b_lbo = (uint32_t)NT * 16u;
        const uint32_t wbuf0_u = smem_u32(wbuf0), mbuf_u = smem_u32(mbuf), bst_u = smem_u32(bst);
        // one group: ntaps (1 or 2) taps of one 32-channel piece; tap t reads window rows shifted by row_off + t * tap_step bytes
        int qc = 0;                                                // partial counter (== c when every group has its own partial)
        // one group: ntaps (1 or 2) taps of one 32-channel piece; tap t reads window rows shifted by row_off + t * tap_step bytes.
        // first / last: the group opens / closes its TMEM partial (always both unless a.gspan).
        auto issue_group = [&](uint32_t a_hi, uint32_t a_lo, uint32_t lbo, uint32_t row_off, uint32_t tap_step, int ntaps, bool first, bool last) {
            const int s = c % S, pb = qc % NPB;
            if (lane == 0) ADEC_TL(2, c);
            mbar_wait(&b_full[s], (c / S) & 1, 300);
            if (lane == 0) ADEC_TL(3, c);
            if (first && qc >= NPB) mbar_wait(&p_empty[pb], ((qc / NPB) - 1) & 1, 400);
            tc_fence_after();
            if (lane == 0) ADEC_TL(4, c);
            const uint32_t bw = bst_u + (uint32_t)s * STAGE_BYTES;
            const uint32_t acc = tmem + (uint32_t)pb * NT;
            if (elect_one()) {
                uint32_t accum = first ? 0u : 1u;
                // descriptor low words; all byte offsets are multiples of 16, so they advance by (bytes >> 4) without touching the LBO field
                const uint32_t ah0 = umma_lo_word(a_hi + row_off, lbo), al0 = umma_lo_word(a_lo + row_off, lbo);
                const uint32_t b0 = umma_lo_word(bw, b_lbo);
                const uint32_t a_ks = (2u * lbo) >> 4, a_t = tap_step >> 4;
                constexpr uint32_t B_KS = (2u * NT * 16u) >> 4, B_T = (uint32_t)TAP_BYTES >> 4, B_PL = (uint32_t)PLANE_B >> 4;
                if (PREC == 3) {
                    // small terms first: A_lo x W_his, A_hi x W_lo
#pragma unroll
                    for (int t = 0; t < 2; ++t)
#pragma unroll
                        for (int ks = 0; ks < KB / 2; ++ks) if (t < ntaps) {
                            umma_f16_lo(acc, al0 + (uint32_t)t * a_t + (uint32_t)ks * a_ks, b0 + (uint32_t)t * B_T + 2u * B_PL + (uint32_t)ks * B_KS, IDESC, accum);
                            accum = 1u;
                        }
#pragma unroll
                    for (int t = 0; t < 2; ++t)
#pragma unroll
                        for (int ks = 0; ks < KB / 2; ++ks)
                            if (t < ntaps) umma_f16_lo(acc, ah0 + (uint32_t)t * a_t + (uint32_t)ks * a_ks, b0 + (uint32_t)t * B_T + B_PL + (uint32_t)ks * B_KS, IDESC, 1u);
                }
#pragma unroll
                for (int t = 0; t < 2; ++t)
#pragma unroll
                    for (int ks = 0; ks < KB / 2; ++ks) if (t < ntaps) {
                        umma_f16_lo(acc, ah0 + (uint32_t)t * a_t + (uint32_t)ks * a_ks, b0 + (uint32_t)t * B_T + (uint32_t)ks * B_KS, IDESC, accum);
                        accum = 1u;
                    }
                umma_commit(&b_empty[s]);
                if (last) umma_commit(&p_full[pb]);
            }
            __syncwarp();
            if (lane == 0) ADEC_TL(5, c);
        };
        const uint32_t lbo1 = (uint32_t)wrp * 16u, lbo2 = (uint32_t)MIDP * 16u;
        const uint32_t tap_step = (uint32_t)a.dil * 16u;
        auto gemm1 = [&]() {        // one tile's conv over its window pieces
            for (int p = 0; p < a.n_pieces; ++p) {
                const int buf = wb;
                if (lane == 0 && mw == 0) ADEC_TL(6, p);
                mbar_wait(&w_full[buf], wround & 1, 200);
                if (lane == 0 && mw == 0) ADEC_TL(7, p);
                if (++wb == a.n_wbuf) { wb = 0; ++wround; }
                const uint32_t a_hi = wbuf0_u + (uint32_t)buf * (uint32_t)win_b;
                const uint32_t a_lo = a_hi + (uint32_t)KB * lbo1;
                for (int t0 = 0; t0 < a.Ktaps; t0 += 2, ++c) {
                    const bool first = !span || t0 == 0, last = !span || t0 + 2 >= a.Ktaps;
                    // span mode: a partial's stages all belong to the warp that owns the partial (qc % NW: the MMAs of one accumulator must come
                    // from one thread, in order).  The other issuer warps wait on every weight stage too and acknowledge it on b_empty (count NW
                    // in this mode), so a stage slot is not refilled before EVERY issuer warp has seen its phase: no waiter can fall two
                    // phases behind (a first version without the acknowledgement dead-locked at batch scale exactly that way)
                    if (span ? qc % NW == mw : c % NW == mw) issue_group(a_hi, a_lo, lbo1, (uint32_t)t0 * tap_step, tap_step, a.Ktaps - t0 >= 2 ? 2 : 1, first, last);
                    else if (span) { mbar_wait(&b_full[c % S], (c / S) & 1, 300); if (lane == 0) mbar_arrive(&b_empty[c % S]); __syncwarp(); }
                    if (last) ++qc;
                }
                if (elect_one()) umma_commit(&w_empty[buf]);
                __syncwarp();
            }
        };
        auto gemm2 = [&]() {        // the fused 1x1 conv over the intermediate pieces
            for (int p = 0; p < NT / CP; ++p, ++mp, ++c) {
                const int mb = mp % MB;
                mbar_wait(&m_full[mb], (mp / MB) & 1, 250);
                const uint32_t m_hi = mbuf_u + (uint32_t)mb * Cfg::MID_BYTES;
                {
                    const bool first = !span || p == 0, last = !span || p == NT / CP - 1;
                    if (span ? qc % NW == mw : c % NW == mw) issue_group(m_hi, m_hi + (uint32_t)KB * lbo2, lbo2, 0u, 0u, 1, first, last);
                    else if (span) { mbar_wait(&b_full[c % S], (c / S) & 1, 300); if (lane == 0) mbar_arrive(&b_empty[c % S]); __syncwarp(); }
                    if (last) ++qc;
                }
                // buffer mb is free for piece mp + MB, which drain group (mp + MB) % NDG writes: signal THAT group's barrier
                if (elect_one()) umma_commit(&m_empty[HALF ? 0 : (mp + MB) % NDG]);
                __syncwarp();
            }
        };
        if (PIPE) {
            if ((int)blockIdx.x < n_tiles) gemm1();
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                if (tile + (int)gridDim.x < n_tiles) gemm1();
                gemm2();
            }
        } else {
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                gemm1();
                if (FUSE) gemm2();
            }
        }
    } else if (warp >= 4 && warp < DRAIN0) {
        // ------------------------------------------------ activation producers: one item = 8 channels (32 B of global) of one window row.
        // The loads of a piece are issued BEFORE the wait for a free window buffer, so global latency overlaps the MMAs that still read
        // the buffer; a.n_wbuf (2..4) buffers let the producers run several pieces ahead.  (Tried and dropped: loading one piece AHEAD
        // into a second register set - 200-500 B of spills per thread, 10.4 -> 17.3 ms per step - and a fully software-pipelined batch
        // sequence with one general row resolver - twice the instructions, 11.3 -> 13.3 ms.  The producers are issue- and register-
        // bound, not load-latency-bound.)
        // Plain (un-fused) launches: the 8 producer warps work as TWO TEAMS of 4 that build alternate window pieces concurrently.  The
        // timeline of the 1x1 / strided / transposed convs showed one piece every ~2600 cycles (a batch of loads, their latency, the
        // conversion) against 400-800 cycles of MMAs per piece: these layers are bound by the producers' load latency, and two pieces in
        // flight hide half of it.  (Safe with any a.n_wbuf >= 2: pieces are consumed in order, so passing the wait for piece q implies every
        // piece <= q - n_wbuf was consumed and no waiter is ever two barrier phases behind.)
        const int pt = tid - 128;
        constexpr bool TEAMS = (!FUSE || ADEC_FUSE_TEAMS) && NPROD == 256;             // kernels that can run two producer teams (a.teams = 1 or 2, ADEC_PLAIN_TEAMS)
        const int nteam = TEAMS ? a.teams : 1, tprod = NPROD / nteam;
        const int team = pt / tprod, ptl = pt - team * tprod;
        int pcnt = 0;                              // running piece counter (team = pcnt % nteam)
        ADEC_TL_DECL(4);
        int wb = 0, wround = 0;                    // window piece counter wp = wround * n_wbuf + wb
        const int RPP = tprod / KB;                // window rows per pass
        constexpr int UNR = (TEAMS || NPROD == 128) ? 5 : 3;   // rows in flight per thread (one 256-bit load each): a pass covers >= 160 rows
        constexpr int UNR_E = NPROD == 128 ? 5 : (TEAMS ? ADEC_UNR_E : 3);   // edge pieces (rare, or every piece of a stacked launch): fewer rows per batch where
                                                                    // the interior batch already takes the registers (3 here spills 120 B instead of 40)
        const int c8 = ptl & (KB - 1), m0 = ptl >> 2;
        const bool halves = a.RG > 1 && a.Cin < 8; // a 4-channel strided conv: the two halves of a block are different x~ rows
        for (TileIter it(blockIdx.x, gridDim.x, n_xtiles, n_ytiles); it.tile < n_tiles; it.next(gridDim.x)) {
            const int xt = it.xt, y = it.y, b = it.b;
            int g = 0, co_tile = y;
            if (a.n_co_tiles != n_ytiles) { g = y / a.n_co_tiles; co_tile = y - g * a.n_co_tiles; }
            const int j0 = xt * TT;
            const float* xg = a.x + (long long)b * a.x_bs + g * a.x_goff;
            const float* sg = a.st_in + (long long)b * a.P * a.st_ld + g * a.st_goff;
            for (int p = 0; p < a.n_pieces; ++p, ++pcnt) {
                const int buf = wb;
                if (pt == 0) ADEC_TL(8, p);
                const uint32_t wpar = (uint32_t)(wround - 1) & 1u;
                bool waited = wround == 0;
                if (++wb == a.n_wbuf) { wb = 0; ++wround; }
                if (TEAMS && nteam > 1 && (pcnt & 1) != team) continue;      // the other team's piece
                unsigned char* hi = wbuf0 + (size_t)buf * win_b + (size_t)c8 * wrp * 16;
                unsigned char* lo = hi + (size_t)KB * wrp * 16;
                const int q = p * CP + c8 * 8;
                int r = 0, ci = q;
                if (a.RG > 1) { r = q >> a.lgCin; ci = q & (a.Cin - 1); }
                const long long i_first = (long long)j0 * a.RG + r;
                const long long i_last = (long long)(j0 + wrows - 1) * a.RG + r;
                if ((a.dbg_flags & 1) && i_first >= a.P && i_last - a.P < a.T) {
                    if (!waited) { mbar_wait(&w_empty[buf], wpar, 500); waited = true; }
                } else if (i_first >= a.P && i_last - a.P < a.T && PRE != ACT_NORM && !halves && !a.stack_L) {
                    // interior piece: every row comes from the chunk
                    const float* xp = xg + ci + (i_first - a.P + (long long)m0 * a.RG) * a.ldx;
                    const long long xstep = (long long)RPP * a.RG * a.ldx;
                    for (int mb = m0; mb < wrows; mb += RPP * UNR, xp += xstep * UNR) {
                        float4 u[UNR], v[UNR];
#pragma unroll
                        for (int k = 0; k < UNR; ++k)
                            if (mb + k * RPP < wrows) ldg256(xp + k * xstep, u[k], v[k]);
                        if (!waited) { mbar_wait(&w_empty[buf], wpar, 500); waited = true; }
#pragma unroll
                        for (int k = 0; k < UNR; ++k) {
                            const int m = mb + k * RPP;
                            if (m < wrows) split_store<PREC>(hi + m * 16, lo + m * 16, apply_act_t<PRE>(u[k], a.slope), apply_act_t<PRE>(v[k], a.slope));
                        }
                    }
                } else {
                    // edge piece: rows from the causal state (stored post-activation), the chunk, or beyond its end (zeros)
                    int ci2 = ci + 4;
                    for (int mb = m0; mb < wrows; mb += RPP * UNR_E) {
                        float4 u[UNR_E], v[UNR_E];
                        unsigned act = 0u;
#pragma unroll
                        for (int k = 0; k < UNR_E; ++k) {
                            const int m = mb + k * RPP;
                            u[k] = v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (m < wrows) {
                                // stacked rows: window row j0 + m of the stack is local row `ml` of stream `sm`
                                int ml = j0 + m;
                                const float* xs = xg;
                                const float* ss = sg;
                                bool live = true;
                                if (a.stack_L) {
                                    const int sm = ml / a.stack_L;
                                    ml -= sm * a.stack_L;
                                    live = sm < a.n_streams;
                                    xs = xg + (long long)sm * a.x_bs;
                                    ss = sg + (long long)sm * a.P * a.st_ld;
                                }
#pragma unroll
                                for (int hf = 0; hf < 2; ++hf) {
                                    int rr = r, cc = ci + 4 * hf;
                                    if (halves) { rr = (q + 4 * hf) >> a.lgCin; cc = (q + 4 * hf) & (a.Cin - 1); if (hf) ci2 = cc; }
                                    const long long i = (long long)ml * a.RG + rr;
                                    long long ti = i - a.P;
                                    if (a.hist_rep && ti < 0) ti = 0;              // non-streaming transposed conv: replicate the first input row
                                    float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
                                    if (live) {
                                        if (ti < 0) w4 = __ldg(reinterpret_cast<const float4*>(ss + i * a.st_ld + cc));
                                        else if (ti < a.T) { w4 = __ldg(reinterpret_cast<const float4*>(xs + ti * a.ldx + cc)); act |= 1u << (2 * k + hf); }
                                    }
                                    if (hf) v[k] = w4; else u[k] = w4;
                                }
                            }
                        }
                        if (!waited) { mbar_wait(&w_empty[buf], wpar, 500); waited = true; }
#pragma unroll
                        for (int k = 0; k < UNR_E; ++k) {
                            const int m = mb + k * RPP;
                            if (m < wrows) {
                                float4 x0 = u[k], x1 = v[k];
                                if ((act >> (2 * k)) & 1u) x0 = PRE == ACT_NORM ? norm4(x0, a.mean + ci, a.scale + ci) : apply_act_t<PRE>(x0, a.slope);
                                if ((act >> (2 * k)) & 2u) x1 = PRE == ACT_NORM ? norm4(x1, a.mean + ci2, a.scale + ci2) : apply_act_t<PRE>(x1, a.slope);
                                split_store<PREC>(hi + m * 16, lo + m * 16, x0, x1);
                            }
                        }
                    }
                }
                fence_async_smem();
                mbar_arrive(&w_full[buf]);
                if (pt == 0) ADEC_TL(9, p);
            }
            // ---- new causal state (conv_layer.py:155): written by the CTA whose tile holds the stream's last output row
            if (co_tile == 0 && g < a.st_groups && a.P > 0) {
                int s_lo = b, s_hi = b - 1;
                if (a.stack_L) {
                    // streams whose last valid row sm * L + Tout - 1 lies in [j0, j0 + TT)
                    s_lo = (j0 - (a.Tout - 1) + a.stack_L - 1) / a.stack_L;
                    if (j0 < a.Tout - 1) s_lo = 0;
                    s_hi = (j0 + TT - 1 - (a.Tout - 1)) / a.stack_L;
                    if (j0 + TT - 1 < a.Tout - 1) s_hi = -1;
                    if (s_hi > a.n_streams - 1) s_hi = a.n_streams - 1;
                } else if (xt == (a.Tout - 1) / TT) {
                    s_hi = b;
                }
                for (int sm = s_lo; sm <= s_hi; ++sm) {
                    const float* xs = a.x + (long long)sm * a.x_bs + g * a.x_goff;
                    const float* ss = a.st_in + (long long)sm * a.P * a.st_ld + g * a.st_goff;
                    float* so = a.st_out + (long long)sm * a.P * a.st_ld + g * a.st_goff;
                    const int nvec = a.P * (a.Cin / 4);
                    for (int idx = pt; idx < nvec; idx += NPROD) {
                        const int r = idx / (a.Cin / 4);
                        const int cc = (idx - r * (a.Cin / 4)) * 4;
                        const long long i = (long long)a.T + r;
                        float4 w4;
                        if (i < a.P) {
                            w4 = *reinterpret_cast<const float4*>(ss + i * a.st_ld + cc);
                        } else {
                            w4 = __ldg(reinterpret_cast<const float4*>(xs + (i - a.P) * a.ldx + cc));
                            if (PRE == ACT_NORM) w4 = norm4(w4, a.mean + cc, a.scale + cc);
                            else w4 = apply_act_t<PRE>(w4, a.slope);
                        }
                        *reinterpret_cast<float4*>(so + (long long)r * a.st_ld + cc) = w4;
                    }
                }
            }
        }
    } else if (warp >= DRAIN0) {
        // ------------------------------------------------ drain warps: register accumulation, fused intermediate, epilogue
        const int dg = (warp - DRAIN0) >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        float racc[NCOL];
        ADEC_TL_DECL(5);
        constexpr bool PREFETCH_RES = FUSE && NT <= 64;     // registers permitting
        float4 rpre[PREFETCH_RES ? PPG : 1][UC / 4];
        int c = 0, mq = 0;
        int m_waits = 0;
        float vmax = 0.f;                                   // largest magnitude this thread produced (fp16-split range check)
        auto drain = [&](float (&acc)[NCOL], int ngroups) {
            for (int gi = 0; gi < ngroups; ++gi, ++c) {
                const int pb = c % NPB;
                if (row == 0 && dg == 0) ADEC_TL(10, c);
                mbar_wait(&p_full[pb], (c / NPB) & 1, 600);
                tc_fence_after();
                if (row == 0 && dg == 0) ADEC_TL(11, c);
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
                if (row == 0 && dg == 0) ADEC_TL(12, c);
            }
        };
        float oacc[PIPE ? NCOL : 1];                         // PIPE: 1x1-conv sums of tile i while racc already holds tile i+1
        if (PIPE && (int)blockIdx.x < n_tiles) {
#pragma unroll
            for (int i = 0; i < NCOL; ++i) racc[i] = 0.f;
            drain(racc, n_g1);                                // conv of the first tile
        }
        for (TileIter it(blockIdx.x, gridDim.x, n_xtiles, n_ytiles); it.tile < n_tiles; it.next(gridDim.x)) {
            const int tile = it.tile, xt = it.xt, y = it.y, b = it.b;
            int g = 0, co_tile = y;
            if (a.n_co_tiles != n_ytiles) { g = y / a.n_co_tiles; co_tile = y - g * a.n_co_tiles; }
            const int j0 = xt * TT;
            // this thread's output row: (stream bo, time t); stacked launches carry several streams per tile and skip the rows between them
            int bo = b, t = j0 + row;
            if (a.stack_L) { bo = t / a.stack_L; t -= bo * a.stack_L; }
            const bool row_ok = t < a.Tout && bo < a.n_streams;
            if (!PIPE) {
#pragma unroll
                for (int i = 0; i < NCOL; ++i) racc[i] = 0.f;
                drain(racc, n_g1);
            }
            if (FUSE) {
                // weight scale out, activation (registers only), so that it overlaps the wait for a free intermediate buffer
#pragma unroll
                for (int i = 0; i < NCOL; i += 4) {
                    const float4 m4 = apply_act_t<PRE>(make_float4(racc[i] * a.w_scale, racc[i + 1] * a.w_scale, racc[i + 2] * a.w_scale, racc[i + 3] * a.w_scale), a.slope);
                    racc[i] = m4.x; racc[i + 1] = m4.y; racc[i + 2] = m4.z; racc[i + 3] = m4.w;
                    vmax = fmaxf(vmax, fmaxf(fmaxf(fabsf(m4.x), fabsf(m4.y)), fmaxf(fabsf(m4.z), fabsf(m4.w))));
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
                        unsigned char* hi = mbuf + mb * Cfg::MID_BYTES;
                        unsigned char* lo = hi + KB * MIDP * 16;
#pragma unroll
                        for (int u = 0; u < UC / 8; ++u) {
                            const int kb = HALF ? dg * (UC / 8) + u : u;       // 16-byte channel block inside the 32-channel piece
                            const float* s8 = racc + pl * UC + u * 8;
                            split_store<PREC>(hi + (kb * MIDP + row) * 16, lo + (kb * MIDP + row) * 16, make_float4(s8[0], s8[1], s8[2], s8[3]),
                                              make_float4(s8[4], s8[5], s8[6], s8[7]));
                        }
                        fence_async_smem();
                        mbar_arrive(&m_full[mb]);
                    }
                }
                mq += NT / CP;
#pragma unroll
                for (int i = 0; i < NCOL; ++i) racc[i] = 0.f;
                if (PREFETCH_RES && a.res && row_ok) {
                    // the skip tensor's rows are known now: fetch them while the MMAs run
#pragma unroll
                    for (int pl = 0; pl < PPG; ++pl) {
                        const int co_l = co_tile * NT + (HALF ? dg * UC : (pl * NDG + dg) * CP);
                        const float* rp = a.res + (long long)bo * a.res_bs + (long long)t * a.ldr + g * a.r_goff + co_l;
#pragma unroll
                        for (int i = 0; i < UC / 4; i += 2) ldg256(rp + 4 * i, rpre[pl][i], rpre[pl][i + 1]);
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
            if (row == 0 && dg == 0) ADEC_TL(13, tile);
            float* const outv = PIPE ? oacc : racc;
            const float oscale = FUSE ? a.w2_scale : a.w_scale;
            // ---- epilogue: row `row` of the tile, this group's PPG pieces of 32 channels
            if (row_ok) {
#pragma unroll
                for (int pl = 0; pl < PPG; ++pl) {
                    const int co_l = co_tile * NT + (HALF ? dg * UC : (pl * NDG + dg) * CP);
                    if (co_l >= a.Cout_g) continue;            // zero-padded part of a channel tile (e.g. 96 outputs in a 128-wide tile)
                    float* v = outv + pl * UC;
#pragma unroll
                    for (int i = 0; i < UC; ++i) v[i] *= oscale;
                    if (a.bias) {
#pragma unroll
                        for (int i = 0; i < UC / 4; ++i) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + g * a.Cout_g + co_l) + i);
                            v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
                        }
                    }
                    if (a.res) {
                        const float* rp = a.res + (long long)bo * a.res_bs + (long long)t * a.ldr + g * a.r_goff + co_l;
                        float4 r4[UC / 4];
#pragma unroll
                        for (int i = 0; i < UC / 4; i += 2) {                                                                      // all loads in flight first
                            if (PREFETCH_RES) { r4[i] = rpre[PREFETCH_RES ? pl : 0][i]; r4[i + 1] = rpre[PREFETCH_RES ? pl : 0][i + 1]; }
                            else ldg256(rp + 4 * i, r4[i], r4[i + 1]);
                        }
#pragma unroll
                        for (int i = 0; i < UC / 4; ++i) {
                            v[4 * i] = r4[i].x + v[4 * i]; v[4 * i + 1] = r4[i].y + v[4 * i + 1];
                            v[4 * i + 2] = r4[i].z + v[4 * i + 2]; v[4 * i + 3] = r4[i].w + v[4 * i + 3];
                        }
                    }
#pragma unroll
                    for (int i = 0; i < UC; ++i) vmax = fmaxf(vmax, fabsf(v[i]));
                    if (a.out_nct) {
                        float* yp = a.y + (long long)bo * a.y_bs + (long long)(g * a.y_goff + co_l) * a.Tout + t;
#pragma unroll
                        for (int i = 0; i < UC; ++i) yp[(long long)i * a.Tout] = v[i];
                    } else {
                        float* yp = a.y + (long long)bo * a.y_bs + (long long)t * a.ldy + g * a.y_goff + co_l;
#pragma unroll
                        for (int i = 0; i < UC / 8; ++i) stg256(yp + 8 * i, v + 8 * i);
                    }
                }
            }
        }
        // every activation is some launch's output: one check here bounds the operands of the next launch's fp16 split
        if (PREC == 3 && a.err && !(vmax < 60000.f)) atomicOr(a.err, 2);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
    if (a.dbg && blockIdx.x == 0 && tid == 0) {
        unsigned long long kt1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(kt1));
        a.dbg[0] = kt0; a.dbg[1] = kt1; a.dbg[2] = (unsigned long long)(clock64() - kc0);
    }
}

}  // namespace adec
