// tcgen05 bring-up probe (sm_100a): D[128 x N] = A_shifted[128 x K] * B[N x K]^T with kind::tf32,
// K-major no-swizzle ("interleave") shared-memory descriptors in the column-block layout
//     smem[(k/4) * ROWS + row][k%4]      (16-byte core-matrix rows, SBO = 128 B, LBO = ROWS*16 B)
// so that a conv tap is just a row-shifted start address.  Checks the result against the CPU and
// prints max error for 1xTF32 and 3xTF32.  build: nvcc -gencode arch=compute_100a,code=sm_100a -o tc_probe tc_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version 1 (Blackwell)
    return d;                 // layout_type 0 = no swizzle
}

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

template <int N, int K, int WROWS, bool SPLIT3>
__global__ void __launch_bounds__(128) probe_kernel(const float* A, const float* B, float* D, int shift) {
    // A: WROWS x K row-major (channels-last window), B: N x K row-major, D: 128 x N
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int LBO_A = WROWS * 16, LBO_B = N * 16;
    float* a_hi = reinterpret_cast<float*>(smem);
    float* a_lo = a_hi + (K / 4) * WROWS * 4;
    float* b_hi = a_lo + (K / 4) * WROWS * 4;
    float* b_lo = b_hi + (K / 4) * N * 4;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    for (int i = tid; i < WROWS * (K / 4); i += 128) {
        const int row = i / (K / 4), c4 = i % (K / 4);
        float4 v = *reinterpret_cast<const float4*>(A + row * K + c4 * 4);
        float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
        float4 l = make_float4(to_tf32(v.x - h.x), to_tf32(v.y - h.y), to_tf32(v.z - h.z), to_tf32(v.w - h.w));
        *reinterpret_cast<float4*>(a_hi + (c4 * WROWS + row) * 4) = h;
        *reinterpret_cast<float4*>(a_lo + (c4 * WROWS + row) * 4) = l;
    }
    for (int i = tid; i < N * (K / 4); i += 128) {
        const int row = i / (K / 4), c4 = i % (K / 4);
        float4 v = *reinterpret_cast<const float4*>(B + row * K + c4 * 4);
        float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
        float4 l = make_float4(to_tf32(v.x - h.x), to_tf32(v.y - h.y), to_tf32(v.z - h.z), to_tf32(v.w - h.w));
        *reinterpret_cast<float4*>(b_hi + (c4 * N + row) * 4) = h;
        *reinterpret_cast<float4*>(b_lo + (c4 * N + row) * 4) = l;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (tensor core)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;

    if (tid == 0) {
        // instruction descriptor: D=f32, A=B=tf32, K-major both, N, M=128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        int first = 1;
        const int npass = SPLIT3 ? 3 : 1;
        for (int pass = 0; pass < npass; ++pass) {
            // small terms first: a_lo*b_hi, a_hi*b_lo, then a_hi*b_hi
            const float* ap = SPLIT3 ? (pass == 0 ? a_lo : a_hi) : a_hi;
            const float* bp = SPLIT3 ? (pass == 1 ? b_lo : b_hi) : b_hi;
            for (int k8 = 0; k8 < K / 8; ++k8) {
                const uint64_t da = make_desc(smem_u32(ap) + k8 * 2 * LBO_A + shift * 16, LBO_A, 128);
                const uint64_t db = make_desc(smem_u32(bp) + k8 * 2 * LBO_B, LBO_B, 128);
                const uint32_t acc = first ? 0u : 1u;
                asm volatile(
                    "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
                    " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                    ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                first = 0;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // wait for the MMAs
    {
        uint32_t ok = 0;
        long long t0 = clock64();
        while (!ok) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
            if (!ok && clock64() - t0 > 2000000000LL) { if (tid == 0) printf("TIMEOUT waiting for MMA\n"); __trap(); }
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w reads TMEM lanes [32w, 32w+32), N columns
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 32; ++i) D[tid * N + c0 + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

template <int N, int K, int WROWS, bool SPLIT3>
int run(int shift) {
    std::vector<float> A(WROWS * K), B(N * K), D(128 * N), R(128 * N);
    srand(1);
    for (auto& v : A) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (auto& v : B) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)A[(m + shift) * K + k] * B[n * K + k];
            R[m * N + n] = (float)s;
        }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    const size_t smem = 2 * (K / 4) * WROWS * 16 + 2 * (K / 4) * N * 16;
    auto kern = probe_kernel<N, K, WROWS, SPLIT3>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<1, 128, smem>>>(dA, dB, dD, shift);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d K=%d shift=%d split3=%d: CUDA error %s\n", N, K, shift, (int)SPLIT3, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (size_t i = 0; i < D.size(); ++i) { maxerr = fmax(maxerr, fabs(D[i] - R[i])); maxref = fmax(maxref, fabs(R[i])); }
    printf("N=%d K=%d WROWS=%d shift=%d split3=%d: max abs err %.3e (max |ref| %.3f)  D[0]=%f R[0]=%f D[last]=%f R[last]=%f\n", N, K, WROWS,
           shift, (int)SPLIT3, maxerr, maxref, D[0], R[0], D.back(), R.back());
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return 0;
}

int main() {
    int rc = 0;
    rc |= run<32, 32, 128, false>(0);
    rc |= run<32, 32, 128, true>(0);
    rc |= run<32, 32, 182, true>(0);
    rc |= run<32, 32, 182, true>(9);
    rc |= run<32, 32, 182, true>(54);
    rc |= run<64, 64, 182, true>(27);
    rc |= run<32, 32, 183, true>(5);     // odd row count -> LBO not a multiple of 128 B
    rc |= run<128, 128, 141, true>(13);
    return rc;
}
