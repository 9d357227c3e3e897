// Shared device/host helpers for librlb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/rlb200.h"

namespace rlb {

constexpr int kNumSMsB200 = 148;

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
struct Timer {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    double ms = 0.0;
    int64_t launches = 0;
    bool pending = false;
};

struct ArenaChunk { char* p; size_t cap; size_t used; };

// exponents (and column sums of squares) of the constant data matrix of a driver call, kept across its passes (ozaki.cu)
struct OzCacheEntry {
    const void* ptr = nullptr;
    int64_t m = 0, n = 0, ld = 0, L = 0;
    int P = 0, elem = 0;
    int* E = nullptr;
    double* ss = nullptr;
    size_t cap_e = 0, cap_s = 0;
    bool valid = false;
};

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int num_sms = kNumSMsB200;
    // workspace arena (device), grown on demand, reused across calls
    void* ws = nullptr;
    size_t ws_bytes = 0;
    // stack allocator for driver-level device buffers (drivers.cu)
    std::vector<ArenaChunk> arena;
    // small pinned host mailbox for return codes / scalars
    void* hbox = nullptr;
    size_t hbox_bytes = 0;
    // row sharding
    int64_t row_offset = 0;
    int64_t m_global = -1;
    rlb200_allreduce_fn allreduce = nullptr;
    void* allreduce_user = nullptr;
    // position of this shard among the row shards (TSQR stacks the k x k factors in this order); -1 = not declared
    int shard_rank = -1, shard_world = 0;
    // native collectives (comm.cu): an NCCL communicator owned by the context, libnccl opened at run time
    void* nccl_comm = nullptr;
    void* nccl_lib = nullptr;
    // tall-GEMM engine of the drivers: 1 = tcgen05 int8 digit slices (ozaki.cu, default), 0 = DMMA (fp64 tensor pipe)
    int fp64_engine = 1;
    // digits per value of the int8-slice engine: 0 = default (6 for fp64 storage, 4 for fp32), else 3..7
    int i8_digits = 0;
    // second stream + events of the int8-slice engine: digit slicing of chunk c+1 runs beside the tensor-core kernel of chunk c
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t aux_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // fork, sliced[2], consumed[2]
    // the matrix a driver declared constant for the duration of a scope (OzConstScope) and its cached exponents
    const void* oz_const_ptr = nullptr;
    OzCacheEntry oz_row, oz_col;
    // the same for the fused engine (ozaki_fused.cu): raw exponents, independent of the digit count
    OzCacheEntry oz2_row, oz2_col;
    // 1 (default): tall products whose shapes allow it run on the fused engine (digits of the tall operand produced inside the
    // tensor-core kernel); 0: always the staged-digit engine (ozaki.cu)
    int i8_fused = 1;
    // per-phase wall-clock times of the last QR driver call (the reference's `times` vectors: rl_cqrrpt.hh:371-384, rl_cqrrt.hh:279-282,
    // rl_bqrrp.hh:582-584), microseconds, recorded only when phase_timing is set (each lap synchronises the stream, like the reference's
    // steady_clock around synchronous BLAS calls)
    bool phase_timing = false;
    bool cqrrpt_orth = false;    // CQRRPT::orthogonalization (rl_cqrrpt.hh:139-142)
    int cqrrpt_qrcp = 0;         // CQRRPT::qrcp (rl_cqrrpt.hh:41): 0 = geqp3 (default), 1 = bqrrp, 2 = hqrrp
    // CQRRPT's HQRRP fields (rl_cqrrpt.hh:134-137; constructor defaults :60-63)
    int64_t cqrrpt_nb_alg = 64, cqrrpt_oversampling = 10;
    int cqrrpt_panel_pivoting = 1, cqrrpt_use_cholqr = 0;
    double bqrrp_tol = 0.0;      // BQRRP::tol (rl_bqrrp.hh:141): 0 = the constructor default, eps of the working type
    std::vector<long long> phase_us;
    // stats
    int64_t launches = 0;
    bool timers_on = false;
    Timer timers[RLB200_TIMER_COUNT];
    std::string err;
};

#define RLB_CUDA_OK(ctx, expr)                                                                      \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + \
                         std::to_string(__LINE__);                                                  \
            return RLB200_ERR_CUDA;                                                                 \
        }                                                                                           \
    } while (0)

#define RLB_REQUIRE(ctx, cond)                                                                      \
    do {                                                                                            \
        if (!(cond)) {                                                                              \
            (ctx)->err = std::string("(" #cond ") was required, but did not hold @") + __FILE__ + ":" + \
                         std::to_string(__LINE__);                                                  \
            return RLB200_ERR_ARG;                                                                  \
        }                                                                                           \
    } while (0)

struct PhaseTimer {
    Ctx* ctx;
    bool on;
    std::chrono::steady_clock::time_point t0, t;
    explicit PhaseTimer(Ctx* c) : ctx(c), on(c->phase_timing) {
        if (on) { cudaStreamSynchronize(c->stream); t0 = t = std::chrono::steady_clock::now(); }
    }
    long long lap() {
        if (!on) return 0;
        cudaStreamSynchronize(ctx->stream);
        const auto n = std::chrono::steady_clock::now();
        const long long us = std::chrono::duration_cast<std::chrono::microseconds>(n - t).count();
        t = n;
        return us;
    }
    long long total() const { return on ? std::chrono::duration_cast<std::chrono::microseconds>(t - t0).count() : 0; }
};

#define RLB_CHECK(expr)                                                                             \
    do {                                                                                            \
        int _rc = (expr);                                                                           \
        if (_rc < 0) return _rc;                                                                    \
    } while (0)

// Reserve `bytes` of device workspace (256-B aligned sub-allocations are carved by the caller).
int ws_reserve(Ctx* ctx, size_t bytes);

struct WsCarver {
    char* base;
    size_t off = 0;
    explicit WsCarver(void* b) : base(static_cast<char*>(b)) {}
    template <typename T>
    T* take(size_t count) {
        T* p = reinterpret_cast<T*>(base + off);
        off += ((count * sizeof(T) + 255) / 256) * 256;
        return p;
    }
};
inline size_t ws_round(size_t bytes) { return ((bytes + 255) / 256) * 256; }

// launch bookkeeping: counts kernels and (optionally) brackets them with CUDA events on ctx->stream.
struct LaunchScope {
    Ctx* ctx;
    int which;
    cudaStream_t stream;
    LaunchScope(Ctx* c, int w, int n_kernels = 1) : LaunchScope(c, w, n_kernels, c->stream) {}
    LaunchScope(Ctx* c, int w, int n_kernels, cudaStream_t st) : ctx(c), which(w), stream(st) {
        ctx->launches += n_kernels;
        Timer& t = ctx->timers[which];
        t.launches += n_kernels;
        if (ctx->timers_on) {
            if (!t.e0) { cudaEventCreate(&t.e0); cudaEventCreate(&t.e1); }
            if (t.pending) { cudaEventSynchronize(t.e1); float ms = 0; cudaEventElapsedTime(&ms, t.e0, t.e1); t.ms += ms; t.pending = false; }
            cudaEventRecord(t.e0, stream);
        }
    }
    ~LaunchScope() {
        Timer& t = ctx->timers[which];
        if (ctx->timers_on) { cudaEventRecord(t.e1, stream); t.pending = true; }
    }
};

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// cp.async (LDGSTS): 16-byte and 8-byte forms with zero-fill when `valid` is false.
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src, bool valid) {
    uint32_t d = smem_u32(smem_dst);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gmem_src, bool valid) {
    uint32_t d = smem_u32(smem_dst);
    int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gmem_src, bool valid) {
    uint32_t d = smem_u32(smem_dst);
    int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// copy a run of `len` (<= E) valid elements (rest zero) of one 16-byte chunk
template <typename T>
__device__ __forceinline__ void load_chunk(T* sdst, const T* gsrc, int valid_elems, bool aligned16) {
    constexpr int E = 16 / sizeof(T);
    if (aligned16 && (valid_elems >= E || valid_elems <= 0)) {
        cp_async_16(sdst, gsrc, valid_elems > 0);
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            bool v = e < valid_elems;
            if (sizeof(T) == 8) cp_async_8(sdst + e, v ? gsrc + e : gsrc, v);
            else                cp_async_4(sdst + e, v ? gsrc + e : gsrc, v);
        }
    }
}

// fp64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  Lane l holds
//   a = A[l/4][l%4], b = B[l%4][l/4], c0/c1 = C[l/4][2*(l%4) + {0,1}].   SASS: DMMA.8x8x4
__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#endif  // __CUDACC__

}  // namespace rlb
