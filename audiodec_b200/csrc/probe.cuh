// mma_probe_kernel: the measured compute ceiling of the conv engine (bench.py's `roofline.compute`).
//
// Every SM runs one CTA whose single issuing thread streams tcgen05.mma (M = 128, N = NT, K = 32 bytes per row: 16 fp16 or 8 tf32)
// from shared-memory operands in the engine's own K-major no-swizzle layout, in groups of 12 accumulating MMAs that rotate over
// four TMEM accumulators - the issue pattern of tc_f16.cuh without producers, drains or global traffic.  What it reports is therefore
// the tensor-pipe + shared-memory-operand ceiling for that tile shape: N = 256 approaches the chip's dense peak, N = 32 shows how far
// the A-operand reads (4 KB per MMA whatever N) hold a narrow layer below it.
#pragma once
#include "tc_kernels.cuh"

namespace adec {

template <int KIND>   // 0 = tf32, 1 = f16
__global__ void __launch_bounds__(128) mma_probe_kernel(int NT, int n_groups, int a_off_rows, int a_pitch_rows, int tap_step_rows, int n_issuers) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    constexpr int KSTEPS = 8;                               // resident K steps, cycled
    unsigned char* as = smem;                               // [KSTEPS * 2 blocks][128 rows][16 B]
    const int a_bytes = (KSTEPS * 2 * a_pitch_rows + a_off_rows + 8 * tap_step_rows + 8) * 16 & ~127;   // A operand: pitch / start row / tap shifts as in the conv windows
    unsigned char* bs = smem + a_bytes;                     // [KSTEPS * 2 blocks][NT rows][16 B]
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < a_bytes / 16 + KSTEPS * 2 * NT; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) { mbar_init(&bar, n_issuers); mbar_fence_init(); }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if ((tid & 31) == 0 && warp < n_issuers) {
        const uint32_t fmt = KIND == 0 ? 2u : 0u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_u = smem_u32(as) + (uint32_t)a_off_rows * 16u, b_u = smem_u32(bs), a_lbo = (uint32_t)a_pitch_rows * 16u, b_lbo = (uint32_t)NT * 16u;
        const int nacc = 512 / NT < 4 ? 512 / NT : 4;
        for (int g = warp; g < n_groups; g += n_issuers) {     // several issuers: groups round robin, each on its own accumulator, unordered
            const uint32_t acc = tmem + (uint32_t)(g % nacc) * NT;
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const uint32_t ks = (uint32_t)((g * 12 + k) % KSTEPS);
                const uint64_t da = umma_desc(a_u + ks * 2u * a_lbo + (uint32_t)((k % 7) * tap_step_rows) * 16u, a_lbo), db = umma_desc(b_u + ks * 2u * b_lbo, b_lbo);
                if (KIND == 0) umma_tf32(acc, da, db, idesc, k ? 1u : 0u);
                else
                    asm volatile(
                        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(acc),
                        "l"(da), "l"(db), "r"(idesc), "r"(k ? 1u : 0u)
                        : "memory");
            }
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0, 900);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

}  // namespace adec
