// Measured int8 tensor-pipe rate of this B200 (the roofline denominator of the digit-slice engine) and the cost of the engine's
// instruction mixes, on operands that stay resident in shared memory (no loads, no epilogue):
//   tcgen05.mma.kind::i8, M = 128 (cta_group::1) or 256 (cta_group::2), K = 32 per instruction, s32 accumulators in TMEM.
// One CTA (or CTA pair) per SM, one issuing thread, at most 8 "K steps" in flight (commit -> mbarrier ring), like the real pipeline.
// Patterns:
//   n256 / n192 / n128 / n64 : 8 instructions of one shape per step                (peak, and the small-N penalty of SS operands)
//   *_sw128                  : the same with SWIZZLE_128B K-major tiles (128-byte rows) instead of the no-swizzle core-matrix layout
//   mix6 / mix7              : the 8 / 10 stacked-digit instructions of one K step of ozaki_fused.cu for 6 / 7 digits (21 / 28 pairs)
//   pairs21                  : 21 separate N = 64 instructions (no stacking)
// usage: peaks_i8 [seconds per sustained run, default 2] [quick]     -> one JSON object on stdout
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/peaks_i8 tools/peaks_i8.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(2); } } while (0)

constexpr int TILE_A = 128 * 32;     // bytes of one digit tile of the M side
constexpr int TILE_B = 64 * 32;      // bytes of one 64-row digit tile of the N side (cta_group::2: each CTA holds half of every instruction's rows)
constexpr int RING = 8;
constexpr int MAXI = 32;

struct Instr { int a_tile, b_tile, ntiles, dcol; };
struct Pattern { int n; int sw128; Instr ins[MAXI]; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {     // SWIZZLE_NONE, K-major: LBO 128 B, SBO 256 B
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint64_t desc_kmajor_sw128(uint32_t saddr) {     // SWIZZLE_128B, K-major: rows of 128 B, SBO 1024 B (8 rows), LBO unused
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ bool wait_bounded(uint32_t bar, uint32_t parity, long long limit = (1ll << 26)) {
    uint32_t done = 0;
    for (long long it = 0; it < limit && !done; ++it)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}

template <int CG, int NI>
__global__ void __launch_bounds__(128, 1) peak_kernel(Pattern pat, long long iters, long long* cycles, int* status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar[RING];
    __shared__ __align__(8) uint64_t bar_done;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    // pseudo-random operand bytes (data-dependent power matters for the sustained figure)
    uint32_t x = 0x9E3779B9u * (blockIdx.x * 128 + tid + 1);
    for (int e = tid; e < (pat.sw128 ? 4 : 1) * (8 * TILE_A + 8 * TILE_B) / 4; e += 128) {
        x = x * 1664525u + 1013904223u;
        reinterpret_cast<uint32_t*>(smem)[e] = x;
    }
    if (tid == 0) {
        for (int s = 0; s < RING; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_done)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    const uint32_t KW = pat.sw128 ? 4 : 1;              // SWIZZLE_128B tiles hold 4 K steps per row
    const uint32_t sa = smem_u32(smem), sb = sa + 8 * TILE_A * KW;
    bool ok = true;
    if (warp == 0 && rank == 0) {
        // The whole warp runs the issue loop (warp-uniform control flow and operands, so the descriptors live in uniform registers);
        // one elected lane issues the MMAs and the commit.
        const uint32_t idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((128 * CG) >> 4) << 24);
        uint32_t da_lo[NI], db_lo[NI], idesc[NI], dcol[NI];
        const uint32_t d_hi = pat.sw128 ? (uint32_t)((1024 >> 4) | (1u << 14) | (2u << 29)) : (uint32_t)((256 >> 4) | (1u << 14));
        const uint32_t lbo = pat.sw128 ? (1u << 16) : ((128u >> 4) << 16);
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const Instr in = pat.ins[i];
            idesc[i] = idesc0 | ((((uint32_t)in.ntiles * 64u) >> 3) << 17);
            da_lo[i] = (((sa + in.a_tile * TILE_A * KW) >> 4) & 0x3FFF) | lbo;
            db_lo[i] = (((sb + in.b_tile * (TILE_B * KW / CG)) >> 4) & 0x3FFF) | lbo;
            dcol[i] = tmem + (uint32_t)in.dcol;
        }
        uint32_t elected = 0;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(elected));
        const long long t0 = clock64();
        for (long long it = 0; it < iters && ok; ++it) {
            const int slot = (int)(it % RING);
            if (it >= RING) ok = wait_bounded(smem_u32(&bar[slot]), (uint32_t)(((it / RING) - 1) & 1));
            const uint32_t acc = it > 0 ? 1u : 0u;
            const uint32_t koff = pat.sw128 ? (uint32_t)((it & 3) * 2) : 0u;       // 32 bytes >> 4 per K step inside the 128-byte row
            if (elected) {
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const uint64_t da = ((uint64_t)d_hi << 32) | (uint64_t)(da_lo[i] + koff);
                    const uint64_t db = ((uint64_t)d_hi << 32) | (uint64_t)(db_lo[i] + koff);
                    if (CG == 1)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
                                     ::"r"(dcol[i]), "l"(da), "l"(db), "r"(idesc[i]), "r"(acc) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
                                     ::"r"(dcol[i]), "l"(da), "l"(db), "r"(idesc[i]), "r"(acc) : "memory");
                }
                if (CG == 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[slot])) : "memory");
                else
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[slot])) : "memory");
            }
            __syncwarp();
        }
        if (elected) {
            if (CG == 1)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_done)) : "memory");
            else
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                             ::"r"(smem_u32(&bar_done)), "h"((uint16_t)3) : "memory");
        }
        __syncwarp();
        if (ok) ok = wait_bounded(smem_u32(&bar_done), 0, 1ll << 31);
        if (tid == 0) { cycles[blockIdx.x] = clock64() - t0; if (!ok) *status = 1; }
    } else if (tid == 0) {
        // the peer of a pair only waits for the final commit
        if (!wait_bounded(smem_u32(&bar_done), 0, 1ll << 31)) *status = 2;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) {
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

static Pattern make_pattern(const std::string& full) {
    Pattern p{};
    std::string name = full;
    const size_t pos = name.find("_sw128");
    if (pos != std::string::npos) { p.sw128 = 1; name = name.substr(0, pos); }
    auto add = [&](int a, int b, int nt, int dc) { p.ins[p.n++] = Instr{a, b, nt, dc}; };
    if (name == "n256") { for (int i = 0; i < 8; ++i) add(i % 6, (i % 2) * 4, 4, (i % 2) * 256); }
    else if (name == "n192") { for (int i = 0; i < 8; ++i) add(i % 6, (i % 2) * 3, 3, (i % 2) * 192); }
    else if (name == "n128") { for (int i = 0; i < 8; ++i) add(i % 6, (i % 4) * 2, 2, (i % 4) * 128); }
    else if (name == "n64") { for (int i = 0; i < 8; ++i) add(i % 6, i, 1, i * 64); }
    else if (name == "n256sameA") { for (int i = 0; i < 8; ++i) add(0, (i % 2) * 4, 4, (i % 2) * 256); }
    else if (name == "n64sameA") { for (int i = 0; i < 8; ++i) add(0, i, 1, i * 64); }
    else if (name == "mix6" || name == "mix7") {
        const int S = name == "mix6" ? 6 : 7;
        for (int s = 0; s < S; ++s)
            for (int t0 = 0; t0 < S - s; t0 += 4) add(s, t0, std::min(4, S - s - t0), (s + t0) * 64);
    } else if (name == "pairs21") {
        for (int s = 0; s < 6; ++s) for (int t = 0; t < 6 - s; ++t) add(s, t, 1, (s + t) * 64);
    }
    return p;
}

int main(int argc, char** argv) {
    const double secs = argc > 1 ? atof(argv[1]) : 2.0;
    const bool quick = argc > 2 && std::string(argv[2]) == "quick";     // only the peak pattern and the two engine mixes, cta_group::1
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int smem = 4 * (8 * TILE_A + 8 * TILE_B) + 8 * 1024;  // 200 KB: one CTA per SM, room for the 128-byte-row tiles
    CK(cudaFuncSetAttribute(peak_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(peak_kernel<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(peak_kernel<1, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(peak_kernel<2, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(peak_kernel<1, 21>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(peak_kernel<2, 21>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    long long* d_cycles; int* d_status;
    CK(cudaMalloc(&d_cycles, sizeof(long long) * sms)); CK(cudaMalloc(&d_status, 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto run = [&](int cg, const Pattern& pat, long long iters, double* ms, double* cyc) -> int {
        CK(cudaMemset(d_status, 0, 4)); CK(cudaMemset(d_cycles, 0, sizeof(long long) * sms));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(sms / cg * cg)); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cg; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaEventRecord(e0));
        if (cg == 1 && pat.n == 8) CK(cudaLaunchKernelEx(&cfg, peak_kernel<1, 8>, pat, iters, d_cycles, d_status));
        else if (cg == 2 && pat.n == 8) CK(cudaLaunchKernelEx(&cfg, peak_kernel<2, 8>, pat, iters, d_cycles, d_status));
        else if (cg == 1 && pat.n == 10) CK(cudaLaunchKernelEx(&cfg, peak_kernel<1, 10>, pat, iters, d_cycles, d_status));
        else if (cg == 2 && pat.n == 10) CK(cudaLaunchKernelEx(&cfg, peak_kernel<2, 10>, pat, iters, d_cycles, d_status));
        else if (cg == 1 && pat.n == 21) CK(cudaLaunchKernelEx(&cfg, peak_kernel<1, 21>, pat, iters, d_cycles, d_status));
        else if (cg == 2 && pat.n == 21) CK(cudaLaunchKernelEx(&cfg, peak_kernel<2, 21>, pat, iters, d_cycles, d_status));
        else { fprintf(stderr, "no kernel for %d instructions\n", pat.n); exit(2); }
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float t; CK(cudaEventElapsedTime(&t, e0, e1));
        int st; CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
        std::vector<long long> h(sms); CK(cudaMemcpy(h.data(), d_cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
        double s = 0; int c = 0;
        for (long long v : h) if (v > 0) { s += (double)v; ++c; }
        *ms = t; *cyc = c ? s / c : 0.0;
        return st;
    };
    const char* names[] = {"n256", "n192", "n128", "n64", "n256sameA", "n64sameA", "mix6", "mix7", "pairs21", "n256_sw128", "n128_sw128", "n64_sw128", "mix6_sw128", "mix7_sw128", "pairs21_sw128"};
    printf("{\"sms\": %d, \"results\": [", sms);
    bool first = true;
    for (int cg = 1; cg <= (quick ? 1 : 2); ++cg) {
        for (const char* nm : names) {
            if (quick && std::string(nm) != "n256" && std::string(nm) != "mix6" && std::string(nm) != "mix7") continue;
            const Pattern pat = make_pattern(nm);
            double units = 0;                      // 64-wide N tiles per K step
            for (int i = 0; i < pat.n; ++i) units += pat.ins[i].ntiles;
            const double macs_step_sm = units * 64.0 * 128.0 * 32.0;       // per SM (a pair does 256 rows on two SMs)
            double ms, cyc;
            // burst: ~10 ms; sustained: `secs` seconds (iteration count from the burst rate)
            long long it_b = 20000;
            int st = run(cg, pat, it_b, &ms, &cyc);
            if (st) { fprintf(stderr, "cta_group::%d %s: status %d (timeout)\n", cg, nm, st); continue; }
            st = run(cg, pat, it_b, &ms, &cyc);
            const double burst = 2.0 * macs_step_sm * it_b * (sms / cg * cg) / (ms * 1e-3) / 1e12;
            const double cyc_step = cyc / it_b;
            const long long it_s = (long long)(secs * 1e3 / ms * it_b);
            double ms2, cyc2;
            const bool do_sus = (std::string(nm) == "n256" || std::string(nm) == "n256_sw128" || std::string(nm) == "mix6" || std::string(nm) == "mix7") && secs > 0;
            double sus = 0, ghz = 0;
            if (do_sus) {
                st = run(cg, pat, it_s, &ms2, &cyc2);
                sus = 2.0 * macs_step_sm * it_s * (sms / cg * cg) / (ms2 * 1e-3) / 1e12;
                ghz = cyc2 / (ms2 * 1e-3) / 1e9;
            }
            printf("%s{\"cta_group\": %d, \"pattern\": \"%s\", \"instr_per_step\": %d, \"n64_units_per_step\": %.0f, \"cycles_per_step\": %.1f, "
                   "\"floor_cycles_per_step\": %.0f, \"burst_tops\": %.1f, \"sustained_tops\": %.1f, \"sustained_ghz\": %.3f}",
                   first ? "" : ", ", cg, nm, pat.n, units, cyc_step, units * 32.0, burst, sus, ghz);
            first = false;
            fflush(stdout);
        }
    }
    printf("]}\n");
    return 0;
}
