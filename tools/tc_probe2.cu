// tcgen05 probe 2: (a) rounding behaviour of the fp32 accumulation in TMEM (bias vs number of accumulation
// steps), (b) MMA issue/throughput for N = 32..256 in SS mode, (c) tcgen05.ld drain time.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)8 << 32) | ((uint64_t)1 << 46);
}
static float tf32r(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; }

// A: 128 x K, B: N x K (already tf32-rounded), column-block layout built in smem; `steps` = K/8 MMAs chained.
template <int N>
__global__ void __launch_bounds__(128) chain_kernel(const float* A, const float* B, float* D, int K, int reps, long long* cyc) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* as = reinterpret_cast<float*>(smem);
    float* bs = as + (K / 4) * 128 * 4;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1)); asm volatile("fence.mbarrier_init.release.cluster;"); }
    for (int i = tid; i < 128 * (K / 4); i += 128) { int row = i / (K / 4), c4 = i % (K / 4); *reinterpret_cast<float4*>(as + (c4 * 128 + row) * 4) = *reinterpret_cast<const float4*>(A + row * K + c4 * 4); }
    for (int i = tid; i < N * (K / 4); i += 128) { int row = i / (K / 4), c4 = i % (K / 4); *reinterpret_cast<float4*>(bs + (c4 * N + row) * 4) = *reinterpret_cast<const float4*>(B + row * K + c4 * 4); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    long long t0 = 0, t1 = 0, t2 = 0;
    uint32_t parity = 0;
    for (int r = 0; r < reps; ++r) {
        if (tid == 0) {
            t0 = clock64();
            for (int k8 = 0; k8 < K / 8; ++k8) {
                const uint64_t da = make_desc(smem_u32(as) + k8 * 2 * 128 * 16, 128 * 16);
                const uint64_t db = make_desc(smem_u32(bs) + k8 * 2 * N * 16, N * 16);
                const uint32_t acc = k8 ? 1u : 0u;
                asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
            t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
        parity ^= 1;
        if (tid == 0) t2 = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    long long t3 = clock64();
    float keep = 0.f;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 32; ++i) { D[tid * N + c0 + i] = __uint_as_float(v[i]); keep += __uint_as_float(v[i]); }
    }
    long long t4 = clock64();
    if (tid == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t0; cyc[2] = t4 - t3; }
    if (keep == 12345.f) D[0] = 0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

template <int N>
void run(int steps, bool positive) {
    const int K = steps * 8;
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    srand(7);
    for (auto& v : A) v = tf32r(positive ? (rand() / (float)RAND_MAX) : (rand() / (float)RAND_MAX - 0.5f) * 2.f);
    for (auto& v : B) v = tf32r(positive ? (rand() / (float)RAND_MAX) : (rand() / (float)RAND_MAX - 0.5f) * 2.f);
    float *dA, *dB, *dD; long long* dC;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dC, 64);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)(K / 4) * (128 + N) * 16;
    if (smem > 200 * 1024) { printf("skip N=%d steps=%d (smem)\n", N, steps); return; }
    auto kern = chain_kernel<N>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<1, 128, smem>>>(dA, dB, dD, K, 3, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d steps=%d: CUDA error %s\n", N, steps, cudaGetErrorString(e)); exit(1); }
    long long c[3];
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(c, dC, 24, cudaMemcpyDeviceToHost);
    double bias = 0, rms = 0; int cnt = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
        double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
        if (fabs(s) > 0.5) { double rel = (D[m * N + n] - s) / fabs(s) * (s > 0 ? 1 : -1); bias += rel; rms += rel * rel; ++cnt; }
    }
    printf("N=%3d steps=%3d %s: signed rel err (toward-zero negative) mean %+.3e rms %.3e over %d | issue %lld clk, issue+complete %lld clk (%.1f clk/MMA), drain(%d cols) %lld clk\n",
           N, steps, positive ? "pos " : "rand", bias / cnt, sqrt(rms / cnt), cnt, c[0], c[1], (double)c[1] / steps, N, c[2]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
}

int main() {
    for (int steps : {1, 4, 16, 64}) { run<32>(steps, false); run<32>(steps, true); }
    run<32>(224, true);
    run<64>(64, true); run<128>(64, true); run<256>(64, true);
    run<64>(16, false); run<128>(16, false); run<256>(16, false);
    return 0;
}
