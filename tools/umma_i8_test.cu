// Step-0 check of the tcgen05 int8 path used by the fp64-by-int8-slices GEMM (randlapack_b200/csrc/ozaki.cu):
// one CTA, D(128 x N, s32 in TMEM) = A(128 x K, s8, K-major) * B(N x K, s8, K-major)^T with no-swizzle core-matrix tiles in shared
// memory, tcgen05.mma.kind::i8, tcgen05.commit -> mbarrier, tcgen05.ld.  Prints PASS/FAIL against a host product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_i8_test tools/umma_i8_test.cu && tools/umma_i8_test
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 64;   // K bytes per tile (2 MMAs of K = 32)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE: ((8,n),2):((1,SBO),LBO) in 16-byte units (cute/atom/mma_traits_sm100.hpp)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // version = 1 (Blackwell)
    return d;                        // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

__global__ void __launch_bounds__(128) umma_test(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ C, int* __restrict__ status) {
    __shared__ __align__(1024) int8_t sA[M * K];
    __shared__ __align__(1024) int8_t sB[N * K];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // fill tiles: element (row, k) -> core (row/8, k/16), row%8, k%16
    for (int e = tid; e < M * K; e += 128) {
        const int row = e / K, k = e % K;
        sA[((row / 8) * (K / 16) + k / 16) * 128 + (row % 8) * 16 + (k % 16)] = A[row * K + k];
    }
    for (int e = tid; e < N * K; e += 128) {
        const int row = e / K, k = e % K;
        sB[((row / 8) * (K / 16) + k / 16) * 128 + (row % 8) * 16 + (k % 16)] = B[row * K + k];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        // instruction descriptor (cute/arch/mma_sm100_desc.hpp): c_format S32 = 2 @4, a/b format INT8 = 1 @7/@10, K-major, N>>3 @17, M>>4 @24
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t lbo = 128, sbo = (K / 16) * 128;
        for (int k = 0; k < K / 32; ++k) {
            const uint64_t da = make_desc(smem_u32(sA) + k * 2 * lbo, lbo, sbo);
            const uint64_t db = make_desc(smem_u32(sB) + k * 2 * lbo, lbo, sbo);
            const uint32_t acc = k > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // wait for the MMAs (bounded spin: a wrong descriptor must not hang the box)
    uint32_t done = 0;
    for (int it = 0; it < (1 << 22) && !done; ++it) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    if (!done) { if (tid == 0) *status = 1; }
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (done) {
        // warp w owns TMEM lanes [32w, 32w+32): thread = row, 8 columns per load
        for (int c0 = 0; c0 < N; c0 += 8) {
            uint32_t v[8];
            const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 8; ++j) C[(warp * 32 + lane) * N + c0 + j] = (int32_t)v[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

int main() {
    std::vector<int8_t> A(M * K), B(N * K);
    srand(1);
    for (auto& v : A) v = (int8_t)(rand() % 129 - 64);
    for (auto& v : B) v = (int8_t)(rand() % 129 - 64);
    std::vector<int32_t> ref(M * N, 0), got(M * N, -1);
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            int32_t s = 0;
            for (int k = 0; k < K; ++k) s += (int32_t)A[i * K + k] * (int32_t)B[j * K + k];
            ref[i * N + j] = s;
        }
    int8_t *dA, *dB; int32_t* dC; int* dS;
    cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dC, got.size() * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
    cudaMemset(dC, 0xFF, got.size() * 4); cudaMemset(dS, 0, 4);
    umma_test<<<1, 128>>>(dA, dB, dC, dS);
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0;
    cudaMemcpy(got.data(), dC, got.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < M * N; ++i) bad += got[i] != ref[i];
    printf("cuda: %s  status %d  mismatches %d / %d  (C[0]=%d ref %d, C[last]=%d ref %d)\n", cudaGetErrorString(e), st, bad, M * N, got[0], ref[0],
           got[M * N - 1], ref[M * N - 1]);
    printf(bad == 0 && st == 0 && e == cudaSuccess ? "UMMA_I8_PASS\n" : "UMMA_I8_FAIL\n");
    return bad != 0;
}
