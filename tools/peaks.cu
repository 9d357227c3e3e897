// Micro-benchmarks for the fp64 pipes of the device (register-resident loops, no memory traffic):
//   DMMA.8x8x4 (mma.sync m8n8k4 f64) and DFMA, with 1..16 independent accumulator chains per warp.
// Prints one JSON object; bench.py / DESIGN.md use "dmma_tflops" as the fp64 tensor-pipe roofline.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <algorithm>

template <int CH>
__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters, double seed) {
    double a = seed + threadIdx.x * 1e-3, b = seed - threadIdx.x * 1e-3;
    double c[CH][2];
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

template <int CH>
__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters, double seed) {
    double a = seed + threadIdx.x * 1e-3, b = seed * 0.5;
    double c[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
static double time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    double best = 1e30;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, (double)ms);
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, 64);
    const int sms = p.multiProcessorCount, iters = 20000;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    double best_dmma = 0, best_dfma = 0;
#define RUN_DMMA(CH, BPS)                                                                                   \
    { int blocks = sms * BPS; double ms = time_ms([&] { dmma_loop<CH><<<blocks, 256>>>(out, iters, 1.0); }); \
      double tf = 2.0 * 256 * CH * (double)iters * 8 * blocks / (ms * 1e-3) / 1e12;                          \
      printf(", \"dmma_ch%d_bps%d\": %.2f", CH, BPS, tf); best_dmma = std::max(best_dmma, tf); }
    RUN_DMMA(1, 1) RUN_DMMA(2, 1) RUN_DMMA(4, 1) RUN_DMMA(8, 1) RUN_DMMA(16, 1) RUN_DMMA(8, 2) RUN_DMMA(16, 2) RUN_DMMA(4, 4)
#define RUN_DFMA(CH, BPS)                                                                                   \
    { int blocks = sms * BPS; double ms = time_ms([&] { dfma_loop<CH><<<blocks, 256>>>(out, iters, 1.0); }); \
      double tf = 2.0 * CH * (double)iters * 256 * blocks / (ms * 1e-3) / 1e12;                              \
      printf(", \"dfma_ch%d_bps%d\": %.2f", CH, BPS, tf); best_dfma = std::max(best_dfma, tf); }
    RUN_DFMA(4, 1) RUN_DFMA(8, 2) RUN_DFMA(16, 4) RUN_DFMA(8, 8)
    printf(", \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f}\n", best_dmma, best_dfma);
    return 0;
}
