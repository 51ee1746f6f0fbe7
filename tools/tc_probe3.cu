// tcgen05 probe 3: error-compensated fp32 GEMM on kind::f16 (two fp16 pieces per operand, three products) against the
// 3xTF32 scheme the conv engine used in round 1, and raw kind::f16 MMA throughput.
//
//   a = A_hi + 2^-11 A_lo          A_hi = fp16(a),          A_lo = fp16((a - A_hi) * 2^11)
//   w * s = W_hi + W_lo            W_hi = fp16(w * s),      W_lo = fp16(w * s - W_hi),      W_his = W_hi * 2^-11   (s = 2^p: max|w| s in [2^12, 2^13))
//   a w s ~= A_lo W_his + A_hi W_lo + A_hi W_hi             (dropped term ~2^-22 relative, like 3xTF32: both formats carry 11 significand bits)
//
// Both kinds use the same K-major no-swizzle column-block layout in BYTES: smem[(kb * ROWS + row) * 16 + byte], kb = 16-byte block
// along K (4 tf32 or 8 fp16), one MMA = 2 blocks (K = 8 tf32 / 16 fp16).  The host prepares the operand images.
// Accumulation is GROUPED like the conv engine: G main MMAs per fresh TMEM partial, partials summed in registers (round to nearest).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe3 tools/tc_probe3.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)8 << 32) | ((uint64_t)1 << 46);
}
template <int KIND> __device__ __forceinline__ void mma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// images: A operands [n_a][kblocks][128 rows][16 B], B operands [n_b][kblocks][N rows][16 B]
// product list (pa, pb): which A image times which B image; per group the products run in list order over the group's K steps.
struct Plan { int n_prod; int pa[3]; int pb[3]; };

template <int KIND, int N>
__global__ void __launch_bounds__(128) gemm_kernel(const unsigned char* Aimg, const unsigned char* Bimg, float* D, int kblocks, int n_a, int n_b,
                                                   Plan plan, int group_steps, float out_scale, int timing_reps, long long* cyc) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* as = smem;
    unsigned char* bs = smem + (size_t)n_a * kblocks * 128 * 16;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1)); asm volatile("fence.mbarrier_init.release.cluster;"); }
    for (int i = tid; i < n_a * kblocks * 128; i += 128) reinterpret_cast<uint4*>(as)[i] = reinterpret_cast<const uint4*>(Aimg)[i];
    for (int i = tid; i < n_b * kblocks * N; i += 128) reinterpret_cast<uint4*>(bs)[i] = reinterpret_cast<const uint4*>(Bimg)[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    // D = f32; A/B format: tf32 = 2 (kind::tf32), f16 = 0, bf16 = 1 (kind::f16)
    const uint32_t fmt = KIND == 0 ? 2u : KIND == 2 ? 1u : 0u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int steps = kblocks / 2;
    const uint32_t a_img = 128u * 16u * kblocks, b_img = (uint32_t)N * 16u * kblocks;
    uint32_t parity = 0;
    float racc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) racc[i] = 0.f;
    long long t_issue = 0, t_done = 0;
    if (timing_reps > 0) {
        // throughput: all steps of product 0 chained into one accumulator, repeated
        long long t0 = clock64();
        for (int r = 0; r < timing_reps; ++r) {
            if (tid == 0) {
                for (int s = 0; s < group_steps; ++s)      // timing: group_steps MMAs cycling over the `steps` K steps held in smem
                    mma<KIND>(tmem, make_desc(smem_u32(as) + (s % steps) * 2 * 128 * 16, 128 * 16), make_desc(smem_u32(bs) + (s % steps) * 2 * N * 16, N * 16), idesc, s ? 1u : 0u);
                if (r == 0) t_issue = clock64() - t0;
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            uint32_t ok = 0;
            while (!ok) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
            parity ^= 1;
        }
        t_done = clock64() - t0;
        if (tid == 0) { cyc[0] = t_issue; cyc[1] = t_done; }
    } else {
        for (int s0 = 0; s0 < steps; s0 += group_steps) {
            const int s1 = s0 + group_steps < steps ? s0 + group_steps : steps;
            if (tid == 0) {
                uint32_t first = 0;
                for (int p = 0; p < plan.n_prod; ++p)
                    for (int s = s0; s < s1; ++s) {
                        mma<KIND>(tmem, make_desc(smem_u32(as) + plan.pa[p] * a_img + s * 2 * 128 * 16, 128 * 16),
                                  make_desc(smem_u32(bs) + plan.pb[p] * b_img + s * 2 * N * 16, N * 16), idesc, first);
                        first = 1;
                    }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            uint32_t ok = 0;
            while (!ok) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
            parity ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c0 = 0; c0 < N; c0 += 16) {
                uint32_t v[16];
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                               "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 16; ++i) racc[c0 + i] = __fadd_rn(racc[c0 + i], __uint_as_float(v[i]));
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < N; ++i) D[tid * N + i] = racc[i] * out_scale;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

static float tf32r(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; }
static uint16_t h16(float x) { __half h = __float2half_rn(x); uint16_t u; memcpy(&u, &h, 2); return u; }
static float f16(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }
static uint16_t b16(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x7FFFu + ((u >> 16) & 1u); return (uint16_t)(u >> 16); }

// place element (row, k) of an operand with `rows` rows into a column-block image; es = element bytes
template <typename T> static void put(std::vector<unsigned char>& img, size_t base, int rows, int row, int k, T v) {
    const int per = 16 / (int)sizeof(T), kb = k / per, e = k % per;
    memcpy(&img[base + ((size_t)kb * rows + row) * 16 + e * sizeof(T)], &v, sizeof(T));
}

static double gauss() { double u = (rand() + 1.0) / (RAND_MAX + 2.0), v = (rand() + 1.0) / (RAND_MAX + 2.0); return sqrt(-2 * log(u)) * cos(6.283185307179586 * v); }

template <int N>
static void run_case(const char* name, int K, double a_scale, int scheme /*0 tf32x3, 1 f16x3, 2 bf16x1, 3 f16x1*/, int group_main) {
    std::vector<float> A(128 * (size_t)K), W((size_t)N * K);
    srand(11);
    for (auto& v : A) { double g = gauss() * 2.0; v = (float)(a_scale * (g > 0 ? g : expm1(g))); }     // ELU-shaped activations
    for (auto& v : W) v = (float)(gauss() * 0.05);
    const int es = scheme == 0 ? 4 : 2, per = 16 / es, kblocks = K / per, kstep = 2 * per;
    int n_a = 1, n_b = 1;
    Plan plan{};
    float out_scale = 1.f;
    std::vector<unsigned char> Ai, Bi;
    if (scheme == 0) {
        n_a = 2; n_b = 2;
        Ai.assign((size_t)n_a * kblocks * 128 * 16, 0); Bi.assign((size_t)n_b * kblocks * N * 16, 0);
        for (int r = 0; r < 128; ++r) for (int k = 0; k < K; ++k) { float h = tf32r(A[(size_t)r * K + k]); put<float>(Ai, 0, 128, r, k, h); put<float>(Ai, (size_t)kblocks * 128 * 16, 128, r, k, tf32r(A[(size_t)r * K + k] - h)); }
        for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) { float h = tf32r(W[(size_t)r * K + k]); put<float>(Bi, 0, N, r, k, h); put<float>(Bi, (size_t)kblocks * N * 16, N, r, k, tf32r(W[(size_t)r * K + k] - h)); }
        plan = Plan{3, {1, 0, 0}, {0, 1, 0}};
    } else if (scheme == 1) {
        n_a = 2; n_b = 3;
        float wmax = 0; for (float v : W) wmax = fmaxf(wmax, fabsf(v));
        int p = 0; while (ldexpf(wmax, p) < 4096.f) ++p; while (ldexpf(wmax, p) >= 8192.f) --p;
        out_scale = ldexpf(1.f, -p);
        Ai.assign((size_t)n_a * kblocks * 128 * 16, 0); Bi.assign((size_t)n_b * kblocks * N * 16, 0);
        for (int r = 0; r < 128; ++r) for (int k = 0; k < K; ++k) {
            const float a = A[(size_t)r * K + k]; const uint16_t hi = h16(a);
            put<uint16_t>(Ai, 0, 128, r, k, hi); put<uint16_t>(Ai, (size_t)kblocks * 128 * 16, 128, r, k, h16((a - f16(hi)) * 2048.f));
        }
        for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) {
            const float w = ldexpf(W[(size_t)r * K + k], p); const uint16_t hi = h16(w);
            put<uint16_t>(Bi, 0, N, r, k, hi); put<uint16_t>(Bi, (size_t)kblocks * N * 16, N, r, k, h16(w - f16(hi)));
            put<uint16_t>(Bi, (size_t)2 * kblocks * N * 16, N, r, k, h16(f16(hi) * (1.f / 2048.f)));
        }
        plan = Plan{3, {1, 0, 0}, {2, 1, 0}};      // A_lo W_his, A_hi W_lo, A_hi W_hi
    } else {
        Ai.assign((size_t)kblocks * 128 * 16, 0); Bi.assign((size_t)kblocks * N * 16, 0);
        for (int r = 0; r < 128; ++r) for (int k = 0; k < K; ++k) put<uint16_t>(Ai, 0, 128, r, k, scheme == 2 ? b16(A[(size_t)r * K + k]) : h16(A[(size_t)r * K + k]));
        for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) put<uint16_t>(Bi, 0, N, r, k, scheme == 2 ? b16(W[(size_t)r * K + k]) : h16(W[(size_t)r * K + k]));
        plan = Plan{1, {0, 0, 0}, {0, 0, 0}};
    }
    unsigned char *dA, *dB; float* dD; long long* dC;
    cudaMalloc(&dA, Ai.size()); cudaMalloc(&dB, Bi.size()); cudaMalloc(&dD, 128 * N * 4); cudaMalloc(&dC, 64);
    cudaMemcpy(dA, Ai.data(), Ai.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, Bi.data(), Bi.size(), cudaMemcpyHostToDevice);
    const size_t smem = Ai.size() + Bi.size();
    if (smem > 220 * 1024) { printf("%s: skipped (smem %zu)\n", name, smem); return; }
    const int steps = K / kstep;
    const int gsteps = group_main > 0 ? group_main : steps;
    cudaError_t e;
    if (scheme == 0) { auto kern = gemm_kernel<0, N>; cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); kern<<<1, 128, smem>>>(dA, dB, dD, kblocks, n_a, n_b, plan, gsteps, out_scale, 0, dC); }
    else if (scheme == 2) { auto kern = gemm_kernel<2, N>; cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); kern<<<1, 128, smem>>>(dA, dB, dD, kblocks, n_a, n_b, plan, gsteps, out_scale, 0, dC); }
    else { auto kern = gemm_kernel<1, N>; cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); kern<<<1, 128, smem>>>(dA, dB, dD, kblocks, n_a, n_b, plan, gsteps, out_scale, 0, dC); }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
    std::vector<float> D(128 * N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double max_abs = 0, ss = 0, se = 0, bias = 0, max_f32 = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
        double s = 0; float sf = 0.f;
        for (int k = 0; k < K; ++k) { s += (double)A[(size_t)m * K + k] * W[(size_t)n * K + k]; sf = fmaf(A[(size_t)m * K + k], W[(size_t)n * K + k], sf); }
        const double err = D[m * N + n] - s;
        max_abs = fmax(max_abs, fabs(err)); ss += s * s; se += err * err; bias += err * (s > 0 ? 1 : -1);
        max_f32 = fmax(max_f32, fabs((double)sf - s));
    }
    const double rms = sqrt(ss / (128 * N));
    printf("%-34s K=%4d N=%3d group=%3d steps | out rms %.3e | max err/rms %.3e  rms err/rms %.3e  signed bias/rms %+.3e | fp32 fma chain max err/rms %.3e\n", name, K, N, gsteps,
           rms, max_abs / rms, sqrt(se / (128 * N)) / rms, bias / (128 * N) / rms, max_f32 / rms);
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
}

template <int KIND, int N>
static void run_timing(const char* name, int steps) {
    const int kblocks = 8 * 2;      // 8 K steps resident, cycled
    std::vector<unsigned char> Ai((size_t)kblocks * 128 * 16, 0), Bi((size_t)kblocks * N * 16, 0);
    unsigned char *dA, *dB; float* dD; long long* dC;
    cudaMalloc(&dA, Ai.size()); cudaMalloc(&dB, Bi.size()); cudaMalloc(&dD, 128 * N * 4); cudaMalloc(&dC, 64);
    cudaMemcpy(dA, Ai.data(), Ai.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, Bi.data(), Bi.size(), cudaMemcpyHostToDevice);
    const size_t smem = Ai.size() + Bi.size();
    auto kern = gemm_kernel<KIND, N>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    Plan plan{1, {0, 0, 0}, {0, 0, 0}};
    const int reps = 8;
    kern<<<1, 128, smem>>>(dA, dB, dD, kblocks, 1, 1, plan, steps, 1.f, reps, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
    long long c[2];
    cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost);
    printf("%-10s N=%3d: %d MMAs x %d reps: issue(first rep) %lld clk, total %lld clk -> %.1f clk/MMA (smem operand bytes/MMA %d)\n", name, N, steps, reps, c[0], c[1],
           (double)c[1] / (steps * reps), 128 * 32 + N * 32);
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
}

int main() {
    // ---- numerics: C=32 k7 conv (K=224), C=128 k7 (K=896); activations O(1) and O(1e-3)
    for (double sc : {1.0, 1e-3, 30.0}) {
        printf("-- activation scale %.0e\n", sc);
        run_case<32>("3xTF32 grouped(4 main MMAs)", 224, sc, 0, 4);
        run_case<32>("fp16x2 3-product grouped(4 main)", 224, sc, 1, 4);
        run_case<32>("fp16x2 3-product grouped(2 main)", 224, sc, 1, 2);
        run_case<32>("3xTF32 one accumulator", 224, sc, 0, 0);
        run_case<32>("fp16x2 3-product one accumulator", 224, sc, 1, 0);
        run_case<64>("3xTF32 grouped(4 main MMAs)", 896, sc, 0, 4);
        run_case<64>("fp16x2 3-product grouped(4 main)", 896, sc, 1, 4);
        run_case<64>("fp16x2 3-product grouped(8 main)", 896, sc, 1, 8);
        run_case<64>("fp16x2 3-product one accumulator", 896, sc, 1, 0);
    }
    printf("-- single-product reference points\n");
    run_case<32>("bf16 x1", 224, 1.0, 2, 0);
    run_case<32>("fp16 x1", 224, 1.0, 3, 0);
    // ---- raw MMA throughput, SS mode, one issuing thread
    run_timing<0, 32>("tf32", 64); run_timing<1, 32>("f16", 64);
    run_timing<0, 64>("tf32", 64); run_timing<1, 64>("f16", 64);
    run_timing<0, 128>("tf32", 64); run_timing<1, 128>("f16", 64);
    run_timing<1, 256>("f16", 32);
    return 0;
}
