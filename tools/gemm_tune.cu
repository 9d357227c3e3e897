// Tile-configuration sweep for the tall-skinny fp64 GEMM kernels (development tool, not part of the product).
// Includes the product kernels verbatim and times each template configuration on the C2-shaped problem.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/gemm_tune tools/gemm_tune.cu
#include "../randlapack_b200/csrc/gemm.cu"
#include <cstdlib>

namespace rlb {
int ws_reserve(Ctx* ctx, size_t bytes) {
    if (bytes <= ctx->ws_bytes) return 0;
    if (ctx->ws) cudaFree(ctx->ws);
    if (cudaMalloc(&ctx->ws, bytes) != cudaSuccess) return RLB200_ERR_ALLOC;
    ctx->ws_bytes = bytes;
    return 0;
}
}  // namespace rlb
using namespace rlb;

__global__ void init_kernel(double* p, size_t n, unsigned seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned x = (unsigned)i * 2654435761u + seed;
        x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
        p[i] = (double)(int)x * (1.0 / 2147483648.0);
    }
}

template <typename F>
static double time_ms(F f, int reps = 3) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, (double)ms);
    }
    return best;
}

int main(int argc, char** argv) {
    const int64_t m = argc > 1 ? atoll(argv[1]) : (1 << 20);
    const int n = 1024, k = 256;
    Ctx ctx;
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    ctx.num_sms = prop.multiProcessorCount;
    double *A, *Om, *Y, *Z;
    cudaMalloc(&A, sizeof(double) * m * n); cudaMalloc(&Om, sizeof(double) * n * k);
    cudaMalloc(&Y, sizeof(double) * m * k); cudaMalloc(&Z, sizeof(double) * n * k);
    init_kernel<<<1184, 256>>>(A, (size_t)m * n, 1); init_kernel<<<148, 256>>>(Om, (size_t)n * k, 2);
    init_kernel<<<1184, 256>>>(Y, (size_t)m * k, 3);
    cudaDeviceSynchronize();
    const double fl = 2.0 * m * n * k, fl_rm = 2.0 * m * k * k;
    printf("{\"m\": %lld", (long long)m);
#define NN(tag, ...)                                                                                                   \
    { double ms = time_ms([&] { launch_nn<double, __VA_ARGS__>(&ctx, m, k, n, 1.0, A, m, Om, n, 0.0, Y, m); });         \
      cudaError_t e = cudaGetLastError();                                                                                \
      printf(", \"nn_%s\": %.2f", tag, e == cudaSuccess ? fl / ms / 1e9 : -1.0); fflush(stdout); }
    //        TM   TN  WGM WGN KS ST MINB
    NN("128x128_w2x4_k32_s3", 128, 128, 2, 4, 32, 3, 1)
    NN("128x128_w2x4_k48_s2", 128, 128, 2, 4, 48, 2, 1)
    NN("128x64_w4x2_k16_s3_b2", 128, 64, 4, 2, 16, 3, 2)
    NN("128x64_w4x2_k32_s2_b2", 128, 64, 4, 2, 32, 2, 2)
    NN("64x64_w2x2_k16_s3_b4", 64, 64, 2, 2, 16, 3, 4)
    NN("64x64_w2x2_k32_s2_b3", 64, 64, 2, 2, 32, 2, 3)
    NN("64x128_w2x4_k16_s3_b2", 64, 128, 2, 4, 16, 3, 2)
    NN("128x128_w4x2_k32_s3", 128, 128, 4, 2, 32, 3, 1)
    NN("64x256_w2x4_k16_s4", 64, 256, 2, 4, 16, 4, 1)
#define RM(tag, ...)                                                                                                   \
    { double ms = time_ms([&] { launch_nn<double, __VA_ARGS__>(&ctx, m, k, k, 1.0, Y, m, Om, k, 0.0, Y, m); });         \
      cudaError_t e = cudaGetLastError();                                                                                \
      printf(", \"rm_%s\": %.2f", tag, e == cudaSuccess ? fl_rm / ms / 1e9 : -1.0); fflush(stdout); }
    RM("64x256_w2x4_k16_s3", 64, 256, 2, 4, 16, 3, 1)
    RM("64x256_w2x4_k16_s4", 64, 256, 2, 4, 16, 4, 1)
    RM("64x256_w4x4_k16_s4", 64, 256, 4, 4, 16, 4, 1)
    RM("64x256_w2x4_k32_s2", 64, 256, 2, 4, 32, 2, 1)
    RM("32x256_w2x4_k16_s3_b2", 32, 256, 2, 4, 16, 3, 2)
#define TN(tag, ...)                                                                                                   \
    { double ms = time_ms([&] { launch_tn<double, __VA_ARGS__>(&ctx, m, n, k, 1.0, A, m, Y, m, 0.0, Z, n, 0); });       \
      cudaError_t e = cudaGetLastError();                                                                                \
      printf(", \"tn_%s\": %.2f", tag, e == cudaSuccess ? fl / ms / 1e9 : -1.0); fflush(stdout); }
    //        T1   T2  WG1 WG2 KS ST MINB
    TN("128x128_w2x4_k32_s3", 128, 128, 2, 4, 32, 3, 1)
    TN("128x128_w2x4_k48_s2", 128, 128, 2, 4, 48, 2, 1)
    TN("128x128_w4x2_k32_s3", 128, 128, 4, 2, 32, 3, 1)
    TN("128x64_w4x2_k16_s3_b2", 128, 64, 4, 2, 16, 3, 2)
    TN("64x64_w2x2_k16_s3_b4", 64, 64, 2, 2, 16, 3, 4)
    TN("64x128_w2x4_k32_s2_b2", 64, 128, 2, 4, 32, 2, 2)
#define SY(tag, ...)                                                                                                   \
    { double ms = time_ms([&] { launch_tn<double, __VA_ARGS__>(&ctx, m, k, k, 1.0, Y, m, Y, m, 0.0, Z, k, 1); });       \
      cudaError_t e = cudaGetLastError();                                                                                \
      printf(", \"syrk_%s\": %.2f", tag, e == cudaSuccess ? 1.0 * m * k * k / ms / 1e9 : -1.0); fflush(stdout); }
    SY("128x128_w2x4_k32_s3", 128, 128, 2, 4, 32, 3, 1)
    SY("64x64_w2x2_k16_s3_b4", 64, 64, 2, 2, 16, 3, 4)
    SY("128x64_w4x2_k16_s3_b2", 128, 64, 4, 2, 16, 3, 2)
    printf("}\n");
    return 0;
}
