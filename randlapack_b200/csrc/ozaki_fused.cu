// Tall products of the drivers on tcgen05.mma.kind::i8 with the digit slicing of the TALL operand fused into the tensor-core kernel.
//
// ozaki.cu stages the digits of both operands in HBM (slicer kernels write them, the MMA kernel streams them back): 8 B read + S B
// written + S B read per value of the 137 GB data matrix and pass, and a slicer class that runs beside the MMA kernel on the same
// shared-memory datapath.  Here the data matrix is read ONCE per pass, as fp64/fp32, by four converter warps of the MMA kernel
// itself: they compute the balanced base-256 digits in registers (one DFMA + integer byte shuffles per value) and store them into the
// pipeline stage the tensor core consumes, in the no-swizzle core-matrix layout.  Only the small operand (Omega, Y: 4x ... 1000x fewer
// values) is still pre-sliced and streamed with bulk copies.
//
// Orientation.  The accumulator of a tile needs S x 64 TMEM columns per 64 indices on the tensor core's N side but nothing per index on
// its M side (lanes), so the SMALL operand sits on the M side (128 of its columns per CTA) and the tall operand on the N side (64 rows
// or columns per CTA, digit tiles stacked):  the tall operand is then shared by only ceil(N / 128) CTAs (2 at k = 256; adjacent in
// launch order, the second read hits L2).
//
//   NN  (rl_rs.hh:153, rl_rf.hh:123):  Y(m x N) = A(m x K) B(K x N)         M side: columns of B (K-major digits of B^T)
//                                                                            N side: rows of A, scale per row, K = columns of A
//   TN  (rl_rs.hh:165, rl_qb.hh:218):  Z(N1 x N2) = A(m x N1)^T Y(m x N2)   M side: columns of Y (MN-major digits)
//                                                                            N side: columns of A, scale per column and 16384-row group
//       + optionally G = Y^T Y (tiles touching the upper triangle) from the same digits of Y (rl_orth.hh:78), on all SD digits, while the
//       main tiles only run the digit pairs of the first sp_main anti-diagonals (6 of 7: 21 instead of 28 pairs).
//
// Output tiles are [M index = lane] x [N index = TMEM column]; the memory layout wants the N index (rows of Y, rows of Z) contiguous, so
// the epilogue transposes through the (by then idle) stage memory and writes 512-byte runs.
// Non-finite inputs: an Inf/NaN entry makes its scaling group's exponent 1025; every output that depends on that group is written as NaN
// (the reference's BLAS path propagates it the same way, potrf then fails and the drivers return their failure codes).
#include "oz_common.cuh"
#include <algorithm>
#include <cstdlib>
#include <vector>

namespace rlb {

constexpr int OZ_RAW_P = -100000;          // "P" that makes oz_exp_from_field return the raw exponent (field - 1022)
constexpr int OZ_NONFINITE_E = 0x7ff - 1022;
constexpr int OZ2_EPI_LD = OZ_BN + 2;      // doubles per M index in the transposing epilogue buffer (528 B: conflict-free 16-byte accesses)

struct Oz2Params {
    // M side: pre-sliced digit tiles, tile (bm, kb, t) at ((bm * nkb_stride + kb) * SD + t) * 4096 (+ g * m_group_stride)
    const int8_t* m_tiles; int64_t m_group_stride; int nkb_stride;
    const int* Em; int64_t em_stride; int rows_m;
    // N side: the tall matrix itself.  NN: element (row, k) at X[row + k * ldx], rows_n rows, kdim columns.
    //                                  TN: element (k, col) at X[g * L + k + col * ldx], rows_n columns, kdim rows in total (over all groups)
    const void* X; int64_t ldx; int64_t rows_n; int64_t kdim; int64_t L;
    const int* En; int64_t en_stride;      // raw exponents (field - 1022): NN per row; TN per (group, column)
    int nkb;                               // K blocks (of 32) per CTA
    int nbm, nbn_main, sp_main;
    void* out; int64_t ldo; int64_t out_group_stride; double alpha, beta;
    int cluster_sync;                      // NN in place (out aliases X): the nbm CTAs of a row tile form a cluster and all finish reading before any writes
    double* gram_out; int64_t gram_group_stride;
    // diagnostics (RLB200_OZ2_DBG bit mask): 1 = per-CTA cycle stamps into dbg, 2 = converters skip the global loads,
    // 4 = no proxy fence, 8 = converters skip the digit arithmetic, 16 = shared digits by plain remote stores + arrives, 32 = no MMAs
    // (timing experiments only: 2/4/8/32 give wrong results)
    long long* dbg; int dbg_flags;
};

__device__ __forceinline__ void oz2_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Shared-memory plan: DST digit stages (M-side digits by bulk copy + N-side digits stored by the converters).  The raw fp64 / fp32 values
// of the tall matrix go global memory -> registers -> digits: staging them in shared memory as well would add 32 KB of shared-memory
// traffic per K step to a main loop that is bound by exactly that (110 KB per K step at 128 B/clk against 672 cycles of MMA issue).
template <int SD, typename T>
struct Oz2Cfg {
    static constexpr int P = OzCfg<SD>::P;
    static constexpr int STAGE_BYTES = OzCfg<SD>::STAGE_BYTES;
    static constexpr int DST_MAX = OZ_SMEM_BUDGET / STAGE_BYTES;
    static constexpr int DST = DST_MAX > 8 ? 8 : DST_MAX;
    static constexpr int SMEM_BYTES = DST * STAGE_BYTES;
    static_assert(DST >= 3, "digit stages");
};

// NG converter groups of 4 warps take the K blocks round-robin (group c: kb = c, c + NG, ...).
// Warp roles: [0, 4 NG) converters (the first 8 also run the epilogue), then one warp each for the M-side bulk copies (lane 0)
// and the MMA issue (lane 0).
template <int NG>
struct Oz2Threads { static constexpr int WARPS = 4 * NG + 2; static constexpr int N = WARPS * 32; };

// streaming read of the tall matrix: read-only path, no L1 allocation (the second CTA that needs the same tile runs on another SM
// and finds it in L2)
__device__ __forceinline__ double oz2_ldg(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float oz2_ldg(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// The tcgen05.mma instructions of one K step: s32 accumulate, signed int8 operands, M = 128; digit s of the M side times the stacked digits
// 0 .. SP-1-s of the N side (<= 4 digit tiles = 256 columns per instruction), accumulator of anti-diagonal d at TMEM column 64 d.
// lo: shared-memory address of the stage's first byte >> 4; the high word of a descriptor (SBO, version) is a constant.
template <int SD, int SP, bool TN>
__device__ __forceinline__ void oz2_issue_one(uint32_t lo, uint32_t tmem, int s, int t0, uint32_t acc) {
    constexpr uint32_t IDESC0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BM >> 4) << 24) | (TN ? ((1u << 15) | (1u << 16)) : 0u);
    constexpr uint64_t HI = ((uint64_t)((TN ? 512 : 256) >> 4) | ((uint64_t)1 << 14)) << 32;
    constexpr uint32_t LBO = (uint32_t)(128 >> 4) << 16;
    // (the start-address field holds bits 4..17 of the shared-memory address: inside a cluster a CTA's window does not start at 0)
    const uint64_t da = HI | (uint64_t)(((lo + (uint32_t)((s * OZ_TILE_A) >> 4)) & 0x3FFFu) | LBO);
    const int nt = (SP - s - t0) < 4 ? (SP - s - t0) : 4;
    const uint32_t idesc = IDESC0 | ((uint32_t)((nt * OZ_BN) >> 3) << 17);
    const uint64_t db = HI | (uint64_t)(((lo + (uint32_t)((SD * OZ_TILE_A + t0 * OZ_TILE_B) >> 4)) & 0x3FFFu) | LBO);
    oz_mma_i8(tmem + (uint32_t)((s + t0) * OZ_BN), da, db, idesc, acc);
}
// Order inside a K step.  The issuing warp pays a fixed cost between two steps (commit, barrier wait, proxy fence: ~150-250 cycles) during
// which the tensor pipe only has what is already queued; the queue is shallow (measured: with the narrow instructions at the end of a step
// the loop runs at MMA time + that cost, 930-1100 cycles per step instead of 790).  So every step but the first ends with its widest
// instructions (N = 256: 128 cycles each): narrow remainders first, then the 4-tile instructions.  The first step of a tile keeps the
// ascending-digit order because its s = 0 instructions initialise the accumulators (accumulate flag 0).
template <int SD, int SP, bool TN>
__device__ __forceinline__ void oz2_issue_step(uint32_t lo, uint32_t tmem, bool first, bool ascending = true) {
    if (first || ascending) {      // (the wide-last sequence below is an experiment: RLB200_OZ2_DBG bit 1024)
#pragma unroll
        for (int s = 0; s < SP; ++s)
#pragma unroll
            for (int t0 = 0; t0 < SP - s; t0 += 4) oz2_issue_one<SD, SP, TN>(lo, tmem, s, t0, (!first || s > 0) ? 1u : 0u);
        return;
    }
    // remainders (fewer than 4 digit tiles), narrowest first
#pragma unroll
    for (int nt = 1; nt < 4; ++nt)
#pragma unroll
        for (int s = SP - 1; s >= 0; --s) {
            const int rem = (SP - s) % 4, t0 = (SP - s) - rem;
            if (rem == nt) oz2_issue_one<SD, SP, TN>(lo, tmem, s, t0, 1u);
        }
    // full 4-tile instructions last
#pragma unroll
    for (int s = SP - 1; s >= 0; --s)
#pragma unroll
        for (int t0 = 0; t0 + 4 <= SP - s; t0 += 4) oz2_issue_one<SD, SP, TN>(lo, tmem, s, t0, 1u);
}

// T: element type of the tall matrix X; TO: element type of `out` (Gram partials are always fp64)
template <int SD, typename T, typename TO, bool TN, int NG>
__global__ void __launch_bounds__(Oz2Threads<NG>::N, 1) oz2_kernel(const Oz2Params p) {
    using Cfg = Oz2Cfg<SD, T>;
    static_assert(NG >= 2, "the epilogue needs 8 warps");
    constexpr int DST = Cfg::DST;
    constexpr int W_PROD = 4 * NG, W_ISSUE = 4 * NG + 1;
    static_assert(OZ_BM * OZ2_EPI_LD * 8 <= DST * Cfg::STAGE_BYTES, "the transposing epilogue buffer must fit in the digit stages");
    constexpr int P = Cfg::P;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bx = (int)(blockIdx.x % (unsigned)p.nbm);
    int by = (int)(blockIdx.x / (unsigned)p.nbm);
    const int g = blockIdx.z;
    const bool gram = by >= p.nbn_main;
    // Cluster of cs = nbm CTAs (launched so when nbm is 2 or 4): the CTAs of one N tile.  They need the same digits of the tall operand, so each
    // converts every cs-th K block and stores its digits into the stage of every CTA of the cluster (distributed shared memory): the global
    // loads, the conversion arithmetic and the load-latency exposure per SM drop by cs.  Gram tiles stream both operands and share nothing.
    uint32_t cs, crank;
    asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(cs));
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    if (gram) {
        by -= p.nbn_main;
        if (by * OZ_BN > bx * OZ_BM + OZ_BM - 1) {             // tiles touching the upper triangle only (row = N index <= column = M index)
            if (cs > 1) { oz_cluster_sync(); oz_cluster_sync(); }      // the two cluster barriers of the CTAs that do run
            return;
        }
    }
    const uint32_t share = gram ? 1u : cs;                     // CTAs that exchange digits with this one
    const uint16_t cmask = (uint16_t)((1u << cs) - 1u);
    long long t_cta0 = 0;
    if (p.dbg) t_cta0 = clock64();
    const int sp = gram ? SD : p.sp_main;                      // anti-diagonals (= leading digits of either operand) this CTA runs
    const int64_t rows_n = gram ? (int64_t)p.rows_m : p.rows_n;
    const int* En = gram ? p.Em + g * p.em_stride : p.En + g * p.en_stride;
    const int nkb = p.nkb;

    extern __shared__ __align__(1024) unsigned char oz2_smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[DST], bar_empty[DST], bar_acc;
    __shared__ uint32_t tmem_base_sh;
    __shared__ int s_en[OZ_BN];                                // effective exponents of the N tile
    __shared__ int s_enmin, s_enmax;
    // K blocks whose stage is free again, published by the producer thread (which waits for EVERY phase of every empty barrier in order).
    // With shared digits a converter visits a given stage only every share * NG blocks - possibly more than DST, i.e. it may skip a phase
    // of that stage's empty barrier, and a parity wait on it would then be ambiguous; it polls this counter instead.
    __shared__ volatile int s_free_upto;

    if (tid == 0) {
        s_free_upto = DST;
        for (int s = 0; s < DST; ++s) {
            // full: the producer's arrive (+ the bytes of its bulk copies) and either the 4 converter warps of this CTA (one arrive each) or,
            // when the digits are shared across a cluster, the bytes of the converters' st.async stores (counted by the same transaction count)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(oz_smem(&bar_full[s])), "r"((gram || (cs > 1 && !(p.dbg_flags & 16))) ? 1 : 5));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(oz_smem(&bar_empty[s])), "r"(share));   // every sharing CTA's MMAs release a slot
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(oz_smem(&bar_acc)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_PROD) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem(&tmem_base_sh)), "n"(OzCfg<SD>::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else if (warp == W_ISSUE) {
        // out-of-range N indices mirror an in-range one so that they do not widen the exponent range of the tile
        const int64_t j0 = (int64_t)by * OZ_BN + lane, j1 = j0 + 32;
        const int e0 = max(En[min(j0, rows_n - 1)], P - 1023), e1 = max(En[min(j1, rows_n - 1)], P - 1023);
        s_en[lane] = e0; s_en[lane + 32] = e1;
        const int mn = __reduce_min_sync(0xffffffffu, min(e0, e1)), mx = __reduce_max_sync(0xffffffffu, max(e0, e1));
        if (lane == 0) { s_enmin = mn; s_enmax = mx; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (cs > 1) oz_cluster_sync();      // every CTA's barriers are initialised before a peer stores digits into it or signals it
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_sh;
    const uint32_t sbase = oz_smem(oz2_smem_raw);

    if (warp < 4 * NG) {
        if (!gram) {
            // ---- converters: 64 N indices x 32 K values per stage, 16 values per thread.  The raw values of K block kb + NG are requested
            // (global -> registers) before block kb is converted, so one block per group is always in flight.
            const int cgp = warp >> 2;        // converter group
            const int cw = warp & 3;          // warp inside the group
            const T* __restrict__ X = reinterpret_cast<const T*>(p.X);
            const T* src;                     // element kk of K block kb: NN src[(kb * 32 + kk) * ldx], TN src[kb * 32 + kk * ldx]
            int sc_hi[TN ? 16 : 1];           // high words of the power-of-two scales 2^(P - E) (low words are zero); 0 = scale 0
            int off;                          // byte offset of this thread's 16-byte chunk inside a digit tile
            // K blocks [0, kb_full) are read without guards, [kb_full, kb_zero) with guards, [kb_zero, nkb) are all zero
            int kb_full, kb_zero;
            int nvalid = 16;                  // TN: valid columns among this thread's 16
            bool tv = true;                   // NN: this thread's row exists
            int64_t kleft = 0;                // TN: K rows from this thread's first row to the end of the matrix
            if constexpr (!TN) {
                const int r = (cw & 1) * 32 + lane, kc = cw >> 1;
                const int64_t row = (int64_t)by * OZ_BN + r;
                tv = row < rows_n;
                src = X + (tv ? row : 0) + (int64_t)(kc * 16) * p.ldx;
                sc_hi[0] = tv ? (P - s_en[r] + 1023) << 20 : 0;
                off = ((r >> 3) * 2 + kc) * 128 + (r & 7) * 16;
                kb_zero = tv ? nkb : 0;
                kb_full = tv ? (int)min((int64_t)nkb, p.kdim / OZ_KB) : 0;
                kleft = p.kdim - kc * 16;     // columns from this thread's first column of block 0 to the end
            } else {
                const int cg = cw;
                const int64_t krow0 = (int64_t)g * p.L + lane;
                const int64_t c0 = (int64_t)by * OZ_BN + cg * 16;
                nvalid = (int)max((int64_t)0, min((int64_t)16, rows_n - c0));
                src = X + krow0 + (nvalid > 0 ? c0 : 0) * p.ldx;
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) sc_hi[kk] = kk < nvalid ? (P - s_en[cg * 16 + kk] + 1023) << 20 : 0;
                off = (cg * 4 + (lane >> 3)) * 128 + (lane & 7) * 16;
                const int64_t left = p.kdim - (int64_t)g * p.L;                   // K rows of this group and beyond
                kleft = left - lane;
                kb_zero = nvalid > 0 ? (int)max((int64_t)0, min((int64_t)nkb, (left + OZ_KB - 1) / OZ_KB)) : 0;
                kb_full = nvalid == 16 ? (int)max((int64_t)0, min((int64_t)nkb, left / OZ_KB)) : 0;
            }
            if (p.dbg_flags & 2) { kb_full = 0; kb_zero = 0; }         // timing experiment: no global traffic for the tall operand
            auto load_block = [&](int kb, T (&raw)[16]) {
                if (kb < kb_full) {
                    if constexpr (!TN) {
                        const T* s0 = src + (int64_t)kb * OZ_KB * p.ldx;
#pragma unroll
                        for (int kk = 0; kk < 16; ++kk) raw[kk] = oz2_ldg(s0 + (int64_t)kk * p.ldx);
                    } else {
                        const T* s0 = src + (int64_t)kb * OZ_KB;
#pragma unroll
                        for (int kk = 0; kk < 16; ++kk) raw[kk] = oz2_ldg(s0 + (int64_t)kk * p.ldx);
                    }
                } else if (kb < kb_zero) {
                    if constexpr (!TN) {
                        const T* s0 = src + (int64_t)kb * OZ_KB * p.ldx;
                        const int64_t cl = kleft - (int64_t)kb * OZ_KB;           // valid columns from this thread's first one
#pragma unroll
                        for (int kk = 0; kk < 16; ++kk) raw[kk] = kk < cl ? oz2_ldg(s0 + (int64_t)kk * p.ldx) : T(0);
                    } else {
                        const T* s0 = src + (int64_t)kb * OZ_KB;
                        const bool kv = (int64_t)kb * OZ_KB < kleft;
#pragma unroll
                        for (int kk = 0; kk < 16; ++kk) raw[kk] = (kv && kk < nvalid) ? oz2_ldg(s0 + (int64_t)kk * p.ldx) : T(0);
                    }
                } else {
#pragma unroll
                    for (int kk = 0; kk < 16; ++kk) raw[kk] = T(0);
                }
            };
            // shared::cluster addresses of the start of every sharing CTA's dynamic shared memory and of its bar_full array
            uint32_t peer_smem[4], peer_full[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t qq = (uint32_t)q < share ? (uint32_t)q : crank;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_smem[q]) : "r"(sbase), "r"(qq));
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_full[q]) : "r"(oz_smem(&bar_full[0])), "r"(qq));
            }
            const int kb_first = (int)crank + (int)share * cgp, kb_stride = (int)share * NG;     // K block kb: CTA kb % share, group (kb / share) % NG
            const bool dbg_on = p.dbg != nullptr && tid == 0;
            long long c_wait = 0, c_conv = 0, c_fence = 0, c_load = 0;
            const bool skip_fence = (p.dbg_flags & 4) != 0, skip_math = (p.dbg_flags & 8) != 0;
            const bool nohint = (p.dbg_flags & 4096) != 0;      // timing experiments: barrier waits without the suspend-time hint
            // L2 prefetch of the tile of K block kb (one 128-byte line per thread: the tile is 64 x 32 elements = 16 KB in fp64): issued
            // several blocks before the register loads, which then find their lines in L2 (~700 cycles) instead of DRAM (> 2000 under load) -
            // the register double buffer alone covers only about one block time (~1200 cycles).  Costs one instruction and no registers.
            const int pf_lines = (int)(OZ_BN * OZ_KB * sizeof(T) / 128);
            const T* pf_src = nullptr;
            {
                const int t = cw * 32 + lane;
                if (t < pf_lines) {
                    if constexpr (!TN) {
                        constexpr int LPC = (int)(OZ_BN * sizeof(T) / 128);              // lines per column of the tile
                        const int64_t row0 = (int64_t)by * OZ_BN + (t % LPC) * (128 / (int)sizeof(T));
                        if (row0 < rows_n) pf_src = X + row0 + (int64_t)(t / LPC) * p.ldx;
                    } else {
                        constexpr int LPC = (int)(OZ_KB * sizeof(T) / 128) > 0 ? (int)(OZ_KB * sizeof(T) / 128) : 1;   // lines per column (32 K rows)
                        const int64_t col = (int64_t)by * OZ_BN + t / LPC;
                        if (col < rows_n) pf_src = X + (int64_t)g * p.L + (t % LPC) * (128 / (int)sizeof(T)) + col * p.ldx;
                    }
                }
            }
            auto prefetch_block = [&](int kb) {
                if (pf_src != nullptr && kb < kb_full) {
                    const T* a = TN ? pf_src + (int64_t)kb * OZ_KB : pf_src + (int64_t)kb * OZ_KB * p.ldx;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
                }
            };
            // convert one K block (values already requested into `raw`) into stage kb % DST and publish it
            auto convert_block = [&](int kb, const T (&raw)[16]) {
                long long t1 = 0, t2 = 0, t3 = 0;
                if (dbg_on) t1 = clock64();
                const int slot = kb % DST;
                if (kb >= DST) {
                    if (share == 1) {
                        if (nohint) oz_mbar_wait(oz_smem(&bar_empty[slot]), (uint32_t)(((kb / DST) - 1) & 1));
                        else oz_mbar_wait_hint(oz_smem(&bar_empty[slot]), (uint32_t)(((kb / DST) - 1) & 1), 20000u);
                    } else {
                        while (s_free_upto <= kb) { if (!nohint) __nanosleep(20); }       // (every poll is a shared-memory access of 8 warps)
                    }
                }
                if (dbg_on) t2 = clock64();
                unsigned char* dst = oz2_smem_raw + slot * Cfg::STAGE_BYTES + SD * OZ_TILE_A + off;
                uint32_t pk[4][SD];
                if (!skip_math) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if constexpr (OzCfg<SD>::P <= 50) {
                            double xv[4], sv[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) { xv[e] = (double)raw[4 * q + e]; sv[e] = __hiloint2double(sc_hi[TN ? 4 * q + e : 0], 0); }
                            oz_convert4<SD>(xv, sv, pk[q]);
                        } else {
                            unsigned long long f[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) f[e] = oz_fixed<SD>((double)raw[4 * q + e], __hiloint2double(sc_hi[TN ? 4 * q + e : 0], 0));
                            oz_pack4<SD>(f, pk[q]);
                        }
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int t = 0; t < SD; ++t) pk[q][t] = (uint32_t)__double_as_longlong((double)raw[4 * q + (t & 3)]);
                }
                if (share == 1) {
                    if (!(p.dbg_flags & 256))      // (256: timing experiment without the digit stores)
#pragma unroll
                    for (int t = 0; t < SD; ++t)
                        if (t < sp) *reinterpret_cast<uint4*>(dst + t * OZ_TILE_B) = make_uint4(pk[0][t], pk[1][t], pk[2][t], pk[3][t]);
                    if (dbg_on) t3 = clock64();
                    // generic-proxy stores -> visible to the tensor core's (async proxy) operand reads: the proxy fence is executed by the
                    // consumer (issuer warp, after its acquire of the full barrier), not by the 4 writer warps of every block - a fence here
                    // costs a MEMBAR per warp and block on the path between "stage free" and "stage full" (p.dbg_flags & 128 restores it)
                    if (p.dbg_flags & 128) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) oz2_arrive(oz_smem(&bar_full[slot]));
                } else {
                    // st.async: each 16-byte store into a CTA of the cluster (this one included) reports its bytes to that CTA's full barrier
                    // (release at cluster scope) - no fence, no arrive and no warp synchronisation on the writer's side; the consumer orders the
                    // stores before the tensor core's async-proxy reads after its acquire (issuer warp)
                    const uint32_t doff = (uint32_t)(slot * Cfg::STAGE_BYTES + SD * OZ_TILE_A + off);
                    if (p.dbg_flags & 16) {
                        // debugging alternative: plain remote stores, writer-side fence, one remote arrive per warp and CTA
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if ((uint32_t)q < share) {
#pragma unroll
                                for (int t = 0; t < SD; ++t)
                                    if (t < sp)
                                        asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};"
                                                     ::"r"(peer_smem[q] + doff + (uint32_t)(t * OZ_TILE_B)), "r"(pk[0][t]), "r"(pk[1][t]), "r"(pk[2][t]), "r"(pk[3][t]) : "memory");
                            }
                        }
                        asm volatile("fence.acq_rel.cluster;" ::: "memory");
                        asm volatile("fence.proxy.async;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if ((uint32_t)q < share)
                                    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(peer_full[q] + (uint32_t)(slot * 8)) : "memory");
                        }
                    } else
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if ((uint32_t)q < share) {
                            const uint32_t bar = peer_full[q] + (uint32_t)(slot * 8);
#pragma unroll
                            for (int t = 0; t < SD; ++t)
                                if (t < sp)
                                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                                                 ::"r"(peer_smem[q] + doff + (uint32_t)(t * OZ_TILE_B)), "r"(pk[0][t]), "r"(pk[1][t]), "r"(pk[2][t]), "r"(pk[3][t]),
                                                   "r"(bar) : "memory");
                        }
                    }
                    if (dbg_on) t3 = clock64();
                }
                if (dbg_on) { c_wait += t2 - t1; c_conv += t3 - t2; c_fence += clock64() - t3; }
            };
            // two register buffers used alternately (no copies: a copy would wait for the loads it is supposed to overlap): while block kb
            // is converted from one buffer the loads of block kb + NG are in flight into the other
            // Order inside an iteration: convert first (the values were requested two blocks of this group ago), THEN request the block
            // after next into the buffer just consumed - the load issue is off the path between "stage free" and "stage full".
            T bufa[16], bufb[16];
            if (kb_first < nkb) load_block(kb_first, bufa);
            if (kb_first + kb_stride < nkb) load_block(kb_first + kb_stride, bufb);
#pragma unroll 1
            constexpr int PF = 5;       // prefetch distance in blocks of this group (the register loads run 2 blocks ahead)
            for (int j = 2; j < PF; ++j) prefetch_block(kb_first + j * kb_stride);
            for (int kb = kb_first; kb < nkb; kb += 2 * kb_stride) {
                long long t0 = 0;
                prefetch_block(kb + PF * kb_stride);
                prefetch_block(kb + (PF + 1) * kb_stride);
                convert_block(kb, bufa);
                if (dbg_on) t0 = clock64();
                if (kb + 2 * kb_stride < nkb) load_block(kb + 2 * kb_stride, bufa);
                if (dbg_on) c_load += clock64() - t0;
                if (kb + kb_stride < nkb) {
                    convert_block(kb + kb_stride, bufb);
                    if (dbg_on) t0 = clock64();
                    if (kb + 3 * kb_stride < nkb) load_block(kb + 3 * kb_stride, bufb);
                    if (dbg_on) c_load += clock64() - t0;
                }
            }
            if (dbg_on) {
                long long* d = p.dbg + ((int64_t)blockIdx.z * gridDim.x + blockIdx.x) * 8;
                d[0] = c_wait; d[1] = c_conv; d[2] = c_fence; d[3] = c_load;
            }
        }
    } else if (tid == W_PROD * 32) {
        // ---- bulk-copy producer: the M-side digits (and, for Gram tiles, the N side from the same array)
        const int8_t* gm = p.m_tiles + g * p.m_group_stride + (int64_t)bx * p.nkb_stride * (SD * OZ_TILE_A);
        const int8_t* gn = p.m_tiles + g * p.m_group_stride + (int64_t)(by >> 1) * p.nkb_stride * (SD * OZ_TILE_A) + (by & 1) * OZ_TILE_B;
        for (int kb = 0; kb < nkb; ++kb) {
            const int slot = kb % DST;
            if (kb >= DST) {
                if (p.dbg_flags & 4096) oz_mbar_wait(oz_smem(&bar_empty[slot]), (uint32_t)(((kb / DST) - 1) & 1));
                else oz_mbar_wait_hint(oz_smem(&bar_empty[slot]), (uint32_t)(((kb / DST) - 1) & 1), 20000u);
                if (share > 1) s_free_upto = kb + 1;      // every sharing CTA's MMAs have consumed block kb - DST: its stage is free in all of them
            }
            const uint32_t bar = oz_smem(&bar_full[slot]);
            const uint32_t dst = sbase + slot * Cfg::STAGE_BYTES;
            // bytes this barrier phase receives: the bulk copies below and, with shared digits, the converters' st.async stores
            const uint32_t bytes = (uint32_t)(sp * OZ_TILE_A + (gram ? SD * OZ_TILE_B : ((share > 1 && !(p.dbg_flags & 16)) ? sp * OZ_TILE_B : 0)));
            if ((p.dbg_flags & 64) && !gram) {     // timing experiment: no M-side bulk copies
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes - (uint32_t)(sp * OZ_TILE_A)) : "memory");
                continue;
            }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            oz_bulk_load(dst, gm + (int64_t)kb * (SD * OZ_TILE_A), (uint32_t)(sp * OZ_TILE_A), bar);
            if (gram) {
                // a 128-column MN-major tile is two 64-column tiles back to back, per digit
#pragma unroll
                for (int t = 0; t < SD; ++t)
                    oz_bulk_load(dst + SD * OZ_TILE_A + t * OZ_TILE_B, gn + (int64_t)kb * (SD * OZ_TILE_A) + t * OZ_TILE_A, OZ_TILE_B, bar);
            }
        }
    } else if (warp == W_ISSUE) {
        // ---- issuer: the whole warp runs the loop (warp-uniform control flow and operands: the descriptors stay in uniform registers and the
        // compiler emits no per-thread serialisation loop around tcgen05.mma), one elected lane issues the MMAs and the commits.
        // Measured (tools/peaks_i8.cu): a single divergent thread that rebuilds both descriptors per instruction issues one MMA per ~110
        // cycles; this form one per <= 45, i.e. the 8 instructions of a 6-digit K step in 790 cycles (floor 672) instead of >= 1000.
        uint32_t elected = 0;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(elected));
        long long c_iwait = 0;
        const uint32_t base_lo0 = sbase >> 4;
        const bool one_stage = (p.dbg_flags & 2048) != 0;      // timing experiments: one stage per issue round
        for (int kb = 0; kb < nkb;) {
            const int slot = kb % DST, slot2 = (kb + 1) % DST;
            long long t0 = 0;
            if (p.dbg) t0 = clock64();
            // up to two stages per issue round (the second only if it is full already): the barrier wait and the proxy fence between two
            // rounds are serial time in which the tensor pipe only has what is queued
            bool two = false;
            if (share == 1) {
                if (p.dbg_flags & 4096) oz_mbar_wait(oz_smem(&bar_full[slot]), (uint32_t)((kb / DST) & 1));
                else oz_mbar_wait_hint(oz_smem(&bar_full[slot]), (uint32_t)((kb / DST) & 1), 20000u);
                if (kb > 0 && kb + 1 < nkb && !one_stage) two = oz_mbar_test(oz_smem(&bar_full[slot2]), (uint32_t)(((kb + 1) / DST) & 1));
                two = __all_sync(0xffffffffu, two);
                if (!gram && !(p.dbg_flags & 512)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the converters' stores of this stage (see there)
            } else {
                // the digits of this stage may come from a peer CTA: acquire at cluster scope, then order those generic-proxy stores before the
                // tensor core's async-proxy reads on the consumer side as well
                if (p.dbg_flags & 4096) oz_mbar_wait_cluster(oz_smem(&bar_full[slot]), (uint32_t)((kb / DST) & 1));
                else oz_mbar_wait_cluster_hint(oz_smem(&bar_full[slot]), (uint32_t)((kb / DST) & 1), 20000u);
                if (kb > 0 && kb + 1 < nkb && !one_stage) two = oz_mbar_test_cluster(oz_smem(&bar_full[slot2]), (uint32_t)(((kb + 1) / DST) & 1));
                two = __all_sync(0xffffffffu, two);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            if (p.dbg) c_iwait += clock64() - t0;
            asm volatile("tcgen05.fence::after_thread_sync;");
            if (elected) {
#pragma unroll 1
                for (int r = 0; r < (two ? 2 : 1); ++r) {
                    const int sl = r ? slot2 : slot;
                    const uint32_t lo = base_lo0 + (uint32_t)sl * (uint32_t)(Cfg::STAGE_BYTES >> 4);
                    if (p.dbg_flags & 32) {}      // timing experiment: no MMAs (stages are released at once)
                    else if (sp == SD) oz2_issue_step<SD, SD, TN>(lo, tmem, kb + r == 0, (p.dbg_flags & 1024) == 0);
                    else oz2_issue_step<SD, (SD > 2 ? SD - 1 : SD), TN>(lo, tmem, kb + r == 0, (p.dbg_flags & 1024) == 0);
                    if (share == 1)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(oz_smem(&bar_empty[sl])) : "memory");
                    else
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                     ::"r"(oz_smem(&bar_empty[sl])), "h"(cmask) : "memory");
                }
            }
            __syncwarp();
            kb += two ? 2 : 1;
        }
        if (elected) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(oz_smem(&bar_acc)) : "memory");
        if (p.dbg && lane == 0) p.dbg[((int64_t)blockIdx.z * gridDim.x + blockIdx.x) * 8 + 4] = c_iwait;
    }
    __syncwarp();

    // ---- epilogue, phase 1: TMEM -> fp64 -> transposing buffer (the stage memory: every read and write of it has completed by now).
    // TMEM lane = M index; warps w and w + 4 share lane quarter w % 4 and take 32 of the 64 N indices each.
    oz_mbar_wait_sleep(oz_smem(&bar_acc), 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    double* epi = reinterpret_cast<double*>(oz2_smem_raw);
    long long t_acc = 0;
    if (p.dbg) t_acc = clock64();
    if (warp < 8) {
        const int quarter = warp & 3, chalf = warp >> 2;
        const int il = quarter * 32 + lane;
        const int64_t im = (int64_t)bx * OZ_BM + il;
        const int em = max(p.Em[g * p.em_stride + min(im, (int64_t)p.rows_m - 1)], P - 1023);
        constexpr int ESHIFT = (2 * P - 16 * (SD - 1)) + 16;
        const bool finite = em != OZ_NONFINITE_E && s_enmax != OZ_NONFINITE_E;
        const bool fast = finite && (em + s_enmin - ESHIFT >= -1022) && (em + s_enmax - ESHIFT <= 1023);
        const int em_hi = (em - ESHIFT + 1023) << 20;
        const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
        for (int c0 = chalf * 32; c0 < chalf * 32 + 32; c0 += 8) {
            uint32_t r[SD][8];
#pragma unroll
            for (int d = 0; d < SD; ++d) {
                if (d < sp) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(r[d][0]), "=r"(r[d][1]), "=r"(r[d][2]), "=r"(r[d][3]), "=r"(r[d][4]), "=r"(r[d][5]), "=r"(r[d][6]), "=r"(r[d][7])
                                 : "r"(lane_addr + (uint32_t)(d * OZ_BN + c0)));
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[d][j] = 0u;
                }
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            double v8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double v = oz_combine<SD>(&r[0][j]);      // 2^16 * sum_d acc_d 256^-d
                const int en = s_en[c0 + j];
                double val;
                if (fast) {
                    val = v * __hiloint2double(em_hi + (en << 20), 0);
                } else if (em == OZ_NONFINITE_E || en == OZ_NONFINITE_E) {
                    val = __longlong_as_double(0x7ff8000000000000ll);
                } else {
                    const int e = em + en - ESHIFT;
                    const int e1 = e / 2, e2 = e - e1;
                    val = (v * oz_pow2(max(-1022, min(1023, e1)))) * oz_pow2(max(-1022, min(1023, e2)));
                }
                v8[j] = val;
            }
            double2* dst = reinterpret_cast<double2*>(epi + il * OZ2_EPI_LD + c0);
#pragma unroll
            for (int j = 0; j < 8; j += 2) dst[j >> 1] = make_double2(v8[j], v8[j + 1]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == W_PROD) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(OzCfg<SD>::TMEM_COLS));
    // in place: every CTA that reads this row tile (the cluster) has consumed all of its stages, i.e. has finished reading the rows
    // and, when digits are shared: no peer still stores into, or signals, this CTA's shared memory once it has passed this barrier
    if (cs > 1) oz_cluster_sync();
    // ---- phase 2: warp = one M index at a time, lanes = 64 consecutive N indices (contiguous in memory)
    {
        TO* outp = reinterpret_cast<TO*>(p.out);
        const int64_t ldo = gram ? (int64_t)p.rows_m : p.ldo;
        const int64_t ogs = gram ? p.gram_group_stride : p.out_group_stride;
        // Gram partials are always fp64
        const int64_t jn = (int64_t)by * OZ_BN + 2 * lane;
        const bool plain = p.alpha == 1.0 && p.beta == 0.0;
#pragma unroll 1
        for (int il = warp; il < OZ_BM; il += Oz2Threads<NG>::WARPS) {
            const int64_t im = (int64_t)bx * OZ_BM + il;
            if (im >= p.rows_m) break;
            const double2 v = *reinterpret_cast<const double2*>(epi + il * OZ2_EPI_LD + 2 * lane);
            if (gram) {
                double* o = p.gram_out + g * ogs + jn + im * ldo;
                if (jn < rows_n) o[0] = plain ? v.x : p.alpha * v.x + (p.beta != 0.0 ? p.beta * o[0] : 0.0);
                if (jn + 1 < rows_n) o[1] = plain ? v.y : p.alpha * v.y + (p.beta != 0.0 ? p.beta * o[1] : 0.0);
            } else {
                TO* o = outp + g * ogs + jn + im * ldo;
                if (jn < rows_n) o[0] = (TO)(plain ? v.x : p.alpha * v.x + (p.beta != 0.0 ? p.beta * (double)o[0] : 0.0));
                if (jn + 1 < rows_n) o[1] = (TO)(plain ? v.y : p.alpha * v.y + (p.beta != 0.0 ? p.beta * (double)o[1] : 0.0));
            }
        }
    }
    if (p.dbg && tid == 32) {
        long long* d = p.dbg + ((int64_t)blockIdx.z * gridDim.x + blockIdx.x) * 8;
        const long long t_end = clock64();
        d[5] = t_acc - t_cta0; d[6] = t_end - t_acc; d[7] = gram ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------
// Persistent A * B kernel (NN): one CTA per SM walks over the output tiles, and the tail of a tile overlaps the head of the next.
//
// oz2_kernel spends 19 k of the 48 k cycles of a K = 1024 tile outside the steady main loop: TMEM allocation, barrier set-up, the DRAM
// latency of the first blocks, the drain of the last DST MMAs (~9 k) and the epilogue (~10.6 k) during which the tensor pipe and the
// converters idle.  Here the barriers' phases, the TMEM allocation and the stage ring run on across tiles: while the last MMAs of tile j
// drain, the converters already fill the ring with the first DST blocks of tile j + 1 (they run ahead of the tensor core by the ring
// depth anyway); then all eight converter warps run phase 1 of tile j's epilogue (TMEM -> fp64 -> a shared-memory buffer of its own),
// release the accumulators, and write tile j to global memory while the tensor core is already consuming the pre-filled ring of tile
// j + 1.  The accumulators are single-buffered (6 x 64 = 384 of 512 TMEM columns), so the MMAs of tile j + 1 wait for phase 1 only.
// ------------------------------------------------------------------------------------------------
struct Oz3Params {
    const int8_t* m_tiles; int nkb_stride;          // M side: digit tiles of B^T, tile (bm, kb, t) at ((bm * nkb_stride + kb) * SD + t) * 4096
    const int* Em; int rows_m;                      // exponents of the columns of B (P applied), number of columns N
    const void* X; int64_t ldx; int64_t rows_n; int64_t kdim;      // the tall matrix A (rows_n x kdim)
    const int* En;                                  // raw row exponents of A
    int nkb, nbm; int64_t ntiles;                   // K blocks per tile, column tiles, tiles in all (tile t: bx = t % nbm, by = t / nbm)
    void* out; int64_t ldo; double alpha, beta;
    int order_asc;                                  // timing experiments: bit 0 (RLB200_OZ3_ONE_STAGE) one stage per issue round, bit 1 (RLB200_OZ3_NO_WAIT_HINT) plain try_wait
};

template <int SD, typename T>
struct Oz3Cfg {
    static constexpr int P = OzCfg<SD>::P;
    static constexpr int STAGE_BYTES = OzCfg<SD>::STAGE_BYTES;
    static constexpr int EPI_BYTES = OZ_BM * OZ2_EPI_LD * 8;
    static constexpr int DST_MAX = (OZ_SMEM_BUDGET - EPI_BYTES) / STAGE_BYTES;
    static constexpr int DST = DST_MAX > 6 ? 6 : DST_MAX;
    static constexpr int SMEM_BYTES = DST * STAGE_BYTES + EPI_BYTES;
    static_assert(DST >= 3, "digit stages");
};

template <int SD, typename T>
__global__ void __launch_bounds__(Oz2Threads<2>::N, 1) oz3_kernel(const Oz3Params p) {
    using Cfg = Oz3Cfg<SD, T>;
    constexpr int NG = 2, DST = Cfg::DST, P = Cfg::P;
    constexpr int W_PROD = 4 * NG, W_ISSUE = 4 * NG + 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb = p.nkb;
    // Tiles of this CTA: ALL column tiles of its row blocks, back to back (row block blockIdx.x + r * gridDim.x, r = 0, 1, ...): the second
    // column tile re-reads the rows of A a few microseconds after the first and finds them in L2.  (With the tiles dealt out one by one
    // the two readers of a row block were different CTAs that drift apart: ncu showed 1.47x the bytes of A read from DRAM.)
    const int64_t nrbt = p.ntiles / p.nbm;
    const int64_t ntl = nrbt > (int64_t)blockIdx.x ? ((nrbt - 1 - blockIdx.x) / gridDim.x + 1) * p.nbm : 0;

    extern __shared__ __align__(1024) unsigned char oz3_smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[DST], bar_empty[DST], bar_acc_full, bar_acc_empty;
    __shared__ uint32_t tmem_base_sh;
    __shared__ int s_en[2][OZ_BN];                  // effective row exponents of the tile, by tile parity
    __shared__ int s_enmin[2], s_enmax[2];
    double* epi = reinterpret_cast<double*>(oz3_smem_raw + DST * Cfg::STAGE_BYTES);

    if (tid == 0) {
        for (int s = 0; s < DST; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 5;" ::"r"(oz_smem(&bar_full[s])));       // producer (+ its bytes) + 4 converter warps
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(oz_smem(&bar_empty[s])));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(oz_smem(&bar_acc_full)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(oz_smem(&bar_acc_empty)));         // the 8 epilogue warps
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_PROD) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem(&tmem_base_sh)), "n"(OzCfg<SD>::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_sh;
    const uint32_t sbase = oz_smem(oz3_smem_raw);

    if (warp < 4 * NG) {
        // ================= converters (+ epilogue) =================
        const int cgp = warp >> 2, cw = warp & 3;
        const T* __restrict__ X = reinterpret_cast<const T*>(p.X);
        const int r = (cw & 1) * 32 + lane, kc = cw >> 1;
        const int off = ((r >> 3) * 2 + kc) * 128 + (r & 7) * 16;
        const int64_t kleft = p.kdim - kc * 16;
        // ---- epilogue of tile (bx, by) whose exponents sit in s_en[par]
        auto epilogue = [&](int bx, int64_t by, int par, uint32_t acc_parity) {
            oz_mbar_wait_sleep(oz_smem(&bar_acc_full), acc_parity);
            asm volatile("tcgen05.fence::after_thread_sync;");
            // all phase-2 reads of the previous tile's buffer are done (and every converter warp has left its conversion loop)
            asm volatile("bar.sync 1, 256;" ::: "memory");
            {
                const int quarter = warp & 3, chalf = warp >> 2;
                const int il = quarter * 32 + lane;
                const int64_t im = (int64_t)bx * OZ_BM + il;
                const int em = max(p.Em[min(im, (int64_t)p.rows_m - 1)], P - 1023);
                constexpr int ESHIFT = (2 * P - 16 * (SD - 1)) + 16;
                const int enmin = s_enmin[par], enmax = s_enmax[par];
                const bool finite = em != OZ_NONFINITE_E && enmax != OZ_NONFINITE_E;
                const bool fast = finite && (em + enmin - ESHIFT >= -1022) && (em + enmax - ESHIFT <= 1023);
                const int em_hi = (em - ESHIFT + 1023) << 20;
                const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
                for (int c0 = chalf * 32; c0 < chalf * 32 + 32; c0 += 8) {
                    uint32_t rr[SD][8];
#pragma unroll
                    for (int d = 0; d < SD; ++d)
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                     : "=r"(rr[d][0]), "=r"(rr[d][1]), "=r"(rr[d][2]), "=r"(rr[d][3]), "=r"(rr[d][4]), "=r"(rr[d][5]), "=r"(rr[d][6]), "=r"(rr[d][7])
                                     : "r"(lane_addr + (uint32_t)(d * OZ_BN + c0)));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    double v8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const double v = oz_combine<SD>(&rr[0][j]);
                        const int en = s_en[par][c0 + j];
                        double val;
                        if (fast) val = v * __hiloint2double(em_hi + (en << 20), 0);
                        else if (em == OZ_NONFINITE_E || en == OZ_NONFINITE_E) val = __longlong_as_double(0x7ff8000000000000ll);
                        else {
                            const int e = em + en - ESHIFT;
                            const int e1 = e / 2, e2 = e - e1;
                            val = (v * oz_pow2(max(-1022, min(1023, e1)))) * oz_pow2(max(-1022, min(1023, e2)));
                        }
                        v8[j] = val;
                    }
                    double2* dst = reinterpret_cast<double2*>(epi + il * OZ2_EPI_LD + c0);
#pragma unroll
                    for (int j = 0; j < 8; j += 2) dst[j >> 1] = make_double2(v8[j], v8[j + 1]);
                }
            }
            // the accumulators are free for the next tile's MMAs
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) oz2_arrive(oz_smem(&bar_acc_empty));
            asm volatile("bar.sync 1, 256;" ::: "memory");      // phase 1 of all eight warps is in the buffer
            {
                T* outp = reinterpret_cast<T*>(p.out);
                const int64_t jn = by * OZ_BN + 2 * lane;
                const bool plain = p.alpha == 1.0 && p.beta == 0.0;
#pragma unroll 1
                for (int il = warp; il < OZ_BM; il += 4 * NG) {
                    const int64_t im = (int64_t)bx * OZ_BM + il;
                    if (im >= p.rows_m) break;
                    const double2 v = *reinterpret_cast<const double2*>(epi + il * OZ2_EPI_LD + 2 * lane);
                    T* o = outp + jn + im * p.ldo;
                    if (jn < p.rows_n) o[0] = (T)(plain ? v.x : p.alpha * v.x + (p.beta != 0.0 ? p.beta * (double)o[0] : 0.0));
                    if (jn + 1 < p.rows_n) o[1] = (T)(plain ? v.y : p.alpha * v.y + (p.beta != 0.0 ? p.beta * (double)o[1] : 0.0));
                }
            }
        };

        T bufa[16], bufb[16];
        int64_t gb0 = 0;                  // blocks of the tiles before this one (ring position)
        int prev_bx = 0; int64_t prev_by = 0;
        const int pre = nkb < DST ? nkb : DST;      // blocks of the next tile converted before the previous tile's epilogue
#pragma unroll 1
        for (int64_t j = 0; j < ntl; ++j) {
            const int bx = (int)(j % p.nbm);
            const int64_t by = (int64_t)blockIdx.x + (j / p.nbm) * gridDim.x;
            const int par = (int)(j & 1);
            // exponents of the tile (out-of-range rows mirror an in-range one so that they do not widen the range)
            if (warp == 0) {
                const int64_t j0 = by * OZ_BN + lane, j1 = j0 + 32;
                const int e0 = max(p.En[min(j0, p.rows_n - 1)], P - 1023), e1 = max(p.En[min(j1, p.rows_n - 1)], P - 1023);
                s_en[par][lane] = e0; s_en[par][lane + 32] = e1;
                const int mn = __reduce_min_sync(0xffffffffu, min(e0, e1)), mx = __reduce_max_sync(0xffffffffu, max(e0, e1));
                if (lane == 0) { s_enmin[par] = mn; s_enmax[par] = mx; }
            }
            const int64_t row = by * OZ_BN + r;
            const bool tv = row < p.rows_n;
            const T* src = X + (tv ? row : 0) + (int64_t)(kc * 16) * p.ldx;
            const int sc_hi = tv ? (P - max(p.En[tv ? row : 0], P - 1023) + 1023) << 20 : 0;
            const int kb_zero = tv ? nkb : 0;
            const int kb_full = tv ? (int)min((int64_t)nkb, p.kdim / OZ_KB) : 0;
            // L2 prefetch: one 128-byte line per thread of the group
            const T* pf_src = nullptr;
            {
                constexpr int LPC = (int)(OZ_BN * sizeof(T) / 128);
                const int tt = cw * 32 + lane;
                if (tt < (int)(OZ_BN * OZ_KB * sizeof(T) / 128)) {
                    const int64_t row0 = by * OZ_BN + (tt % LPC) * (128 / (int)sizeof(T));
                    if (row0 < p.rows_n) pf_src = X + row0 + (int64_t)(tt / LPC) * p.ldx;
                }
            }
            // (a bulk form - one cp.async.bulk.prefetch.L2 of 512 bytes per column, issued by one warp - keeps the 128 line requests per block
            //  off the L1 data pipe but is far slower: A*Omega 14.8 ms against 10.8-11.3 ms on the same box)
            auto prefetch_block = [&](int kb) {
                if (pf_src != nullptr && kb < kb_full) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf_src + (int64_t)kb * OZ_KB * p.ldx));
            };
            auto load_block = [&](int kb, T (&raw)[16]) {
                const T* s0 = src + (int64_t)kb * OZ_KB * p.ldx;
                if (kb < kb_full) {
#pragma unroll
                    for (int kk = 0; kk < 16; ++kk) raw[kk] = oz2_ldg(s0 + (int64_t)kk * p.ldx);
                } else if (kb < kb_zero) {
                    const int64_t cl = kleft - (int64_t)kb * OZ_KB;
#pragma unroll
                    for (int kk = 0; kk < 16; ++kk) raw[kk] = kk < cl ? oz2_ldg(s0 + (int64_t)kk * p.ldx) : T(0);
                } else {
#pragma unroll
                    for (int kk = 0; kk < 16; ++kk) raw[kk] = T(0);
                }
            };
            auto convert_block = [&](int kb, const T (&raw)[16]) {
                const int64_t gb = gb0 + kb;
                const int slot = (int)(gb % DST);
                if (gb >= DST) { if (p.order_asc & 2) oz_mbar_wait(oz_smem(&bar_empty[slot]), (uint32_t)(((gb / DST) - 1) & 1)); else oz_mbar_wait_hint(oz_smem(&bar_empty[slot]), (uint32_t)(((gb / DST) - 1) & 1), 20000u); }
                unsigned char* dst = oz3_smem_raw + slot * Cfg::STAGE_BYTES + SD * OZ_TILE_A + off;
                uint32_t pk[4][SD];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if constexpr (OzCfg<SD>::P <= 50) {
                        double xv[4], sv[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) { xv[e] = (double)raw[4 * q + e]; sv[e] = __hiloint2double(sc_hi, 0); }
                        oz_convert4<SD>(xv, sv, pk[q]);
                    } else {
                        unsigned long long f[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) f[e] = oz_fixed<SD>((double)raw[4 * q + e], __hiloint2double(sc_hi, 0));
                        oz_pack4<SD>(f, pk[q]);
                    }
                }
#pragma unroll
                for (int t2 = 0; t2 < SD; ++t2) *reinterpret_cast<uint4*>(dst + t2 * OZ_TILE_B) = make_uint4(pk[0][t2], pk[1][t2], pk[2][t2], pk[3][t2]);
                __syncwarp();
                if (lane == 0) oz2_arrive(oz_smem(&bar_full[slot]));
            };
            constexpr int PF = 5;
            for (int q = 2; q < PF; ++q) prefetch_block(cgp + q * NG);
            if (cgp < nkb) load_block(cgp, bufa);
            if (cgp + NG < nkb) load_block(cgp + NG, bufb);
            bool epi_done = j == 0;       // the previous tile's epilogue runs once the first `pre` blocks of this tile are in the ring
#pragma unroll 1
            for (int kb = cgp; kb < nkb; kb += 2 * NG) {
                if (!epi_done && kb >= pre) { epilogue(prev_bx, prev_by, par ^ 1, (uint32_t)((j - 1) & 1)); epi_done = true; }
                prefetch_block(kb + PF * NG);
                prefetch_block(kb + (PF + 1) * NG);
                convert_block(kb, bufa);
                if (kb + 2 * NG < nkb) load_block(kb + 2 * NG, bufa);
                if (kb + NG < nkb) {
                    if (!epi_done && kb + NG >= pre) { epilogue(prev_bx, prev_by, par ^ 1, (uint32_t)((j - 1) & 1)); epi_done = true; }
                    convert_block(kb + NG, bufb);
                    if (kb + 3 * NG < nkb) load_block(kb + 3 * NG, bufb);
                }
            }
            if (!epi_done) epilogue(prev_bx, prev_by, par ^ 1, (uint32_t)((j - 1) & 1));
            gb0 += nkb;
            prev_bx = bx; prev_by = by;
        }
        if (ntl > 0) epilogue(prev_bx, prev_by, (int)((ntl - 1) & 1), (uint32_t)((ntl - 1) & 1));
    } else if (tid == W_PROD * 32) {
        // ================= M-side bulk copies =================
        int64_t gb = 0;
        for (int64_t j = 0; j < ntl; ++j) {
            const int bx = (int)(j % p.nbm);
            const int8_t* gm = p.m_tiles + (int64_t)bx * p.nkb_stride * (SD * OZ_TILE_A);
            for (int kb = 0; kb < nkb; ++kb, ++gb) {
                const int slot = (int)(gb % DST);
                if (gb >= DST) { if (p.order_asc & 2) oz_mbar_wait(oz_smem(&bar_empty[slot]), (uint32_t)(((gb / DST) - 1) & 1)); else oz_mbar_wait_hint(oz_smem(&bar_empty[slot]), (uint32_t)(((gb / DST) - 1) & 1), 20000u); }
                const uint32_t bar = oz_smem(&bar_full[slot]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(SD * OZ_TILE_A)) : "memory");
                oz_bulk_load(sbase + slot * Cfg::STAGE_BYTES, gm + (int64_t)kb * (SD * OZ_TILE_A), (uint32_t)(SD * OZ_TILE_A), bar);
            }
        }
    } else if (warp == W_ISSUE) {
        // ================= MMA issue =================
        uint32_t elected = 0;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(elected));
        const uint32_t base_lo0 = sbase >> 4;
        int64_t gb = 0;
        for (int64_t j = 0; j < ntl; ++j) {
            for (int kb = 0; kb < nkb;) {
                const int slot = (int)(gb % DST);
                if (p.order_asc & 2) oz_mbar_wait(oz_smem(&bar_full[slot]), (uint32_t)((gb / DST) & 1)); else oz_mbar_wait_hint(oz_smem(&bar_full[slot]), (uint32_t)((gb / DST) & 1), 20000u);
                // the previous tile's accumulators have been read
                if (kb == 0 && j > 0) oz_mbar_wait(oz_smem(&bar_acc_empty), (uint32_t)((j - 1) & 1));
                // If the next stage is full already, both are issued behind one barrier wait / proxy fence: between two issue rounds the
                // tensor pipe only has what is queued (measured: ~135 cycles per round on top of the 790 of the MMAs of a K step)
                const int slot2 = (int)((gb + 1) % DST);
                bool two = (kb + 1 < nkb) && !(p.order_asc & 1) && oz_mbar_test(oz_smem(&bar_full[slot2]), (uint32_t)(((gb + 1) / DST) & 1));
                two = __all_sync(0xffffffffu, two);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the converters' generic-proxy stores of this stage
                asm volatile("tcgen05.fence::after_thread_sync;");
                if (elected) {
                    const uint32_t lo = base_lo0 + (uint32_t)slot * (uint32_t)(Cfg::STAGE_BYTES >> 4);
                    oz2_issue_step<SD, SD, false>(lo, tmem, kb == 0, true);
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(oz_smem(&bar_empty[slot])) : "memory");
                    if (two) {
                        const uint32_t lo2 = base_lo0 + (uint32_t)slot2 * (uint32_t)(Cfg::STAGE_BYTES >> 4);
                        oz2_issue_step<SD, SD, false>(lo2, tmem, false, true);
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(oz_smem(&bar_empty[slot2])) : "memory");
                    }
                    if (kb + (two ? 1 : 0) == nkb - 1)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(oz_smem(&bar_acc_full)) : "memory");
                }
                __syncwarp();
                kb += two ? 2 : 1;
                gb += two ? 2 : 1;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == W_PROD) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(OzCfg<SD>::TMEM_COLS));
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
constexpr int OZ2_NG = 2;      // converter groups (4 warps each): 2 fit 168 registers per thread without spills and measured fastest
template <int SD, typename T, typename TO, bool TN>
static int oz2_configure(Ctx* ctx) {
    static bool done_dev[64] = {};
    bool& done = done_dev[ctx->device & 63];
    if (!done) {
        RLB_CUDA_OK(ctx, cudaFuncSetAttribute(oz2_kernel<SD, T, TO, TN, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Oz2Cfg<SD, T>::SMEM_BYTES));
        RLB_CUDA_OK(ctx, cudaFuncSetAttribute(oz2_kernel<SD, T, TO, TN, OZ2_NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, Oz2Cfg<SD, T>::SMEM_BYTES));
        done = true;
    }
    return 0;
}

static bool oz2_cache_hit(const OzCacheEntry& e, const void* ptr, int64_t m, int64_t n, int64_t ld, int64_t L, int elem) {
    return e.valid && e.ptr == ptr && e.m == m && e.n == n && e.ld == ld && e.L == L && e.P == OZ_RAW_P && e.elem == elem;
}
static int oz2_cache_reserve(Ctx* ctx, OzCacheEntry& e, size_t n_e, size_t n_s) {
    if (n_e > e.cap_e || n_s > e.cap_s) {
        RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->aux_stream) RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->aux_stream));
        if (n_e > e.cap_e) {
            if (e.E) cudaFree(e.E);
            e.E = nullptr; e.cap_e = 0;
            if (cudaMalloc(&e.E, n_e * sizeof(int)) != cudaSuccess) { cudaGetLastError(); ctx->err = "exponent cache allocation failed"; return RLB200_ERR_ALLOC; }
            e.cap_e = n_e;
        }
        if (n_s > e.cap_s) {
            if (e.ss) cudaFree(e.ss);
            e.ss = nullptr; e.cap_s = 0;
            if (cudaMalloc(&e.ss, n_s * sizeof(double)) != cudaSuccess) { cudaGetLastError(); ctx->err = "exponent cache allocation failed"; return RLB200_ERR_ALLOC; }
            e.cap_s = n_s;
        }
    }
    e.valid = false;
    return 0;
}
static void oz2_cache_set(OzCacheEntry& e, const void* ptr, int64_t m, int64_t n, int64_t ld, int64_t L, int elem) {
    e.ptr = ptr; e.m = m; e.n = n; e.ld = ld; e.L = L; e.P = OZ_RAW_P; e.elem = elem; e.valid = true;
}
// Rows per accumulation chunk of the long-contraction (A^T Y) launches = rows per exponent group of the N side: a CTA runs one chunk of
// one tile.  OZ_CHUNK = 16384 is the exactness bound of the int32 accumulators: an anti-diagonal sums up to 7 digit pairs x L rows of
// products <= 2^14, and 7 * 2^14 * 2^14 < 2^31.  Longer chunks are faster on random data (m = 2^21, n = 1024, k = 256: 13.5 ms at 16384 rows,
// 12.3-12.8 ms at 32768: fewer epilogues and partial sums) but can overflow on adversarial digits, so they stay an experiment
// (RLB200_OZ2_CHUNK) and are never the default.
static int64_t oz2_chunk_rows(int64_t m) {
    static const int64_t chunk_env = getenv("RLB200_OZ2_CHUNK") ? atoll(getenv("RLB200_OZ2_CHUNK")) : 0;
    const int64_t chunk = chunk_env > 0 ? chunk_env : OZ_CHUNK;
    return std::min<int64_t>(chunk, ((m + OZ_KB - 1) / OZ_KB) * OZ_KB);
}

// Raw row exponents, raw column-group exponents and the sums of squares of the constant data matrix of a driver scope, in one sweep
// (both caches are independent of the digit count: the kernels apply the floor P - 1023 themselves).
template <typename T>
static int oz2_fill_const_cache(Ctx* ctx, const T* A, int64_t m, int64_t K, int64_t lda, cudaStream_t st) {
    const int64_t L = oz2_chunk_rows(m);
    const int64_t nchunks = (m + L - 1) / L, ncta = (m + 255) / 256;
    const bool fused = (L % 256 == 0) && K <= 8192 && (size_t)nchunks * K >= (size_t)ncta;
    RLB_CHECK(oz2_cache_reserve(ctx, ctx->oz2_row, (size_t)m, 0));
    RLB_CHECK(oz2_cache_reserve(ctx, ctx->oz2_col, (size_t)nchunks * K, (size_t)nchunks * K));
    if (fused) {
        LaunchScope ls(ctx, RLB200_TIMER_I8_SLICE, 2, st);
        RLB_CUDA_OK(ctx, cudaMemsetAsync(ctx->oz2_col.E, 0, sizeof(int) * (size_t)nchunks * K, st));
        RLB_CUDA_OK(ctx, cudaMemsetAsync(ctx->oz2_col.ss, 0, sizeof(double) * (size_t)nchunks * K, st));
        oz_rowcolexp_kernel<T><<<(unsigned)ncta, 256, (size_t)K * sizeof(int), st>>>(A, lda, m, (int)K, L, OZ_RAW_P, ctx->oz2_row.E, ctx->oz2_col.E, ctx->oz2_col.ss);
        oz_field_to_exp_kernel<<<(unsigned)((nchunks * K + 255) / 256), 256, 0, st>>>(ctx->oz2_col.E, nchunks * K, OZ_RAW_P);
    } else {
        LaunchScope ls(ctx, RLB200_TIMER_I8_SLICE, 2, st);
        oz_rowexp_kernel<T><<<(unsigned)((m + 255) / 256), 256, 0, st>>>(A, lda, m, (int)K, OZ_RAW_P, ctx->oz2_row.E);
        oz_colexp_kernel<T><<<(unsigned)((K * nchunks + 7) / 8), 256, 0, st>>>(A, lda, m, (int)K, L, (int)nchunks, OZ_RAW_P, ctx->oz2_col.E, ctx->oz2_col.ss);
    }
    RLB_CUDA_OK(ctx, cudaGetLastError());
    oz2_cache_set(ctx->oz2_row, A, m, K, lda, 0, (int)sizeof(T));
    oz2_cache_set(ctx->oz2_col, A, m, K, lda, L, (int)sizeof(T));
    return 0;
}
void oz2_cache_destroy(Ctx* ctx) {
    for (OzCacheEntry* e : {&ctx->oz2_row, &ctx->oz2_col}) {
        if (e->E) cudaFree(e->E);
        if (e->ss) cudaFree(e->ss);
        *e = OzCacheEntry();
    }
}

template <int SD, typename T, typename TO, bool TN>
static int oz2_launch(Ctx* ctx, dim3 grid, cudaStream_t st, const Oz2Params& p_in) {
    Oz2Params p = p_in;
    static const int dbg_flags = getenv("RLB200_OZ2_DBG") ? atoi(getenv("RLB200_OZ2_DBG")) : 0;
    p.dbg = nullptr; p.dbg_flags = dbg_flags;
    const int64_t nctas = (int64_t)grid.x * grid.z;
    if (dbg_flags & 1) { cudaMalloc(&p.dbg, (size_t)nctas * 64); cudaMemsetAsync(p.dbg, 0, (size_t)nctas * 64, st); }
    static const int ng = getenv("RLB200_OZ2_NG") ? atoi(getenv("RLB200_OZ2_NG")) : OZ2_NG;
    // Clusters (digits of the tall operand shared by the nbm CTAs of an N tile): measured on B200 (m = 2^21, n = 1024, k = 256, 2 converter
    // groups) TN 14.2 ms shared / 14.9 ms not shared, NN 13.8 / 12.6 - so TN shares, NN does not unless it is in place (where the cluster
    // barrier is what makes it correct).  RLB200_OZ2_SHARE=0/1 forces either for experiments.
    const int csz = (p.nbm == 2 || p.nbm == 4) ? p.nbm : 1;
    if (p.cluster_sync && csz == 1 && p.nbm != 1) { ctx->err = "in-place product needs 1, 2 or 4 column tiles"; return RLB200_ERR_ARG; }
    static const int share_env = getenv("RLB200_OZ2_SHARE") ? atoi(getenv("RLB200_OZ2_SHARE")) : -1;
    const bool share = share_env >= 0 ? share_env != 0 : TN;
    if (csz > 1 && (share || p.cluster_sync)) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid; cfg.blockDim = dim3(Oz2Threads<OZ2_NG>::N); cfg.dynamicSmemBytes = Oz2Cfg<SD, T>::SMEM_BYTES; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        RLB_CUDA_OK(ctx, cudaLaunchKernelEx(&cfg, oz2_kernel<SD, T, TO, TN, OZ2_NG>, p));
    } else if (ng == 3) oz2_kernel<SD, T, TO, TN, 3><<<grid, Oz2Threads<3>::N, Oz2Cfg<SD, T>::SMEM_BYTES, st>>>(p);
    else oz2_kernel<SD, T, TO, TN, OZ2_NG><<<grid, Oz2Threads<OZ2_NG>::N, Oz2Cfg<SD, T>::SMEM_BYTES, st>>>(p);
    RLB_CUDA_OK(ctx, cudaGetLastError());
    if (p.dbg) {
        cudaStreamSynchronize(st);
        std::vector<long long> h((size_t)nctas * 8);
        cudaMemcpy(h.data(), p.dbg, h.size() * 8, cudaMemcpyDeviceToHost);
        double a[7] = {0, 0, 0, 0, 0, 0, 0};
        int64_t cnt = 0;
        for (int64_t i = 0; i < nctas; ++i) {
            if (h[(size_t)i * 8 + 7] != 0 || h[(size_t)i * 8 + 5] == 0) continue;
            for (int j = 0; j < 7; ++j) a[j] += (double)h[(size_t)i * 8 + j];
            ++cnt;
        }
        if (cnt)
            fprintf(stderr, "[oz2 dbg] %s SD=%d nkb=%d: %lld main CTAs; cycles/CTA: converter{wait_empty %.0f, convert %.0f, fence+arrive %.0f, "
                            "load issue %.0f}, issuer wait_full %.0f, start->accumulators %.0f, epilogue %.0f\n",
                    TN ? "TN" : "NN", SD, p.nkb, (long long)cnt, a[0] / cnt, a[1] / cnt, a[2] / cnt, a[3] / cnt, a[4] / cnt, a[5] / cnt, a[6] / cnt);
        cudaFree(p.dbg);
    }
    return 0;
}

// C(m x N) = alpha * A(m x K) * B(K x N) + beta * C; C may be A itself (same pointer and leading dimension, N == K, beta == 0)
template <int SD, typename T>
static int oz2_nn(Ctx* ctx, int64_t m, int64_t N, int64_t K, double alpha, const T* A, int64_t lda, const T* B, int64_t ldb, double beta, T* C, int64_t ldc) {
    using Cfg = OzCfg<SD>;
    RLB_CHECK((oz2_configure<SD, T, T, false>(ctx)));
    const int nkb = (int)((K + OZ_KB - 1) / OZ_KB);
    const int nbm = (int)((N + OZ_BM - 1) / OZ_BM);
    const int64_t nbn = (m + OZ_BN - 1) / OZ_BN;
    RLB_REQUIRE(ctx, nbn * nbm < ((int64_t)1 << 31));
    ArenaScope as(ctx);
    int* Eb = as.take<int>((size_t)N); if (!Eb) return RLB200_ERR_ALLOC;
    int8_t* bt = as.take<int8_t>((size_t)nbm * nkb * SD * OZ_TILE_A); if (!bt) return RLB200_ERR_ALLOC;
    cudaStream_t st = ctx->stream;
    const int* Ea = nullptr;
    if (ctx->oz_const_ptr == (const void*)A) {
        if (!oz2_cache_hit(ctx->oz2_row, A, m, K, lda, 0, (int)sizeof(T))) RLB_CHECK(oz2_fill_const_cache<T>(ctx, A, m, K, lda, st));
        Ea = ctx->oz2_row.E;
    } else {
        int* e = as.take<int>((size_t)m); if (!e) return RLB200_ERR_ALLOC;
        LaunchScope ls(ctx, RLB200_TIMER_I8_SLICE, 1);
        oz_rowexp_kernel<T><<<(unsigned)((m + 255) / 256), 256, 0, st>>>(A, lda, m, (int)K, OZ_RAW_P, e);
        RLB_CUDA_OK(ctx, cudaGetLastError());
        Ea = e;
    }
    {
        LaunchScope ls(ctx, RLB200_TIMER_I8_SLICE, 2);
        oz_colexp_kernel<T><<<(unsigned)((N + 7) / 8), 256, 0, st>>>(B, ldb, K, (int)N, K, 1, Cfg::P, Eb, nullptr);
        oz_slice_cols_kernel<SD, OZ_BM, T><<<dim3(nbm, (nkb + 7) / 8), 128, 0, st>>>(B, ldb, K, (int)N, nkb, Eb, bt);
        RLB_CUDA_OK(ctx, cudaGetLastError());
    }
    // persistent kernel (the tail of a tile overlaps the head of the next) unless the product is in place (that one needs the cluster barrier
    // between the last read and the first write of a row tile) - RLB200_OZ3=0 falls back to one CTA per tile
    static const bool use_oz3 = !(getenv("RLB200_OZ3") && atoi(getenv("RLB200_OZ3")) == 0);
    if (use_oz3 && (const void*)A != (const void*)C) {
        static bool done_dev[64] = {};
        bool& done = done_dev[ctx->device & 63];
        if (!done) {
            RLB_CUDA_OK(ctx, cudaFuncSetAttribute(oz3_kernel<SD, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, Oz3Cfg<SD, T>::SMEM_BYTES));
            done = true;
        }
        Oz3Params q{};
        q.m_tiles = bt; q.nkb_stride = nkb; q.Em = Eb; q.rows_m = (int)N;
        q.X = A; q.ldx = lda; q.rows_n = m; q.kdim = K; q.En = Ea;
        q.nkb = nkb; q.nbm = nbm; q.ntiles = nbn * nbm;
        q.out = C; q.ldo = ldc; q.alpha = alpha; q.beta = beta;
        static const int order_asc = (getenv("RLB200_OZ3_ONE_STAGE") ? 1 : 0) | (getenv("RLB200_OZ3_NO_WAIT_HINT") ? 2 : 0);
        q.order_asc = order_asc;
        LaunchScope ls(ctx, RLB200_TIMER_I8_MMA_NN);
        const unsigned grid = (unsigned)std::min<int64_t>(nbn, (int64_t)ctx->num_sms);
        oz3_kernel<SD, T><<<grid, Oz2Threads<2>::N, Oz3Cfg<SD, T>::SMEM_BYTES, st>>>(q);
        RLB_CUDA_OK(ctx, cudaGetLastError());
        return 0;
    }
    Oz2Params p{};
    p.m_tiles = bt; p.m_group_stride = 0; p.nkb_stride = nkb;
    p.Em = Eb; p.em_stride = 0; p.rows_m = (int)N;
    p.X = A; p.ldx = lda; p.rows_n = m; p.kdim = K; p.L = 0;
    p.En = Ea; p.en_stride = 0;
    p.nkb = nkb; p.nbm = nbm; p.nbn_main = (int)nbn; p.sp_main = SD;
    p.out = C; p.ldo = ldc; p.out_group_stride = 0; p.alpha = alpha; p.beta = beta;
    p.gram_out = nullptr; p.gram_group_stride = 0;
    p.cluster_sync = ((const void*)A == (const void*)C) ? 1 : 0;
    LaunchScope ls(ctx, RLB200_TIMER_I8_MMA_NN);
    return oz2_launch<SD, T, T, false>(ctx, dim3((unsigned)(nbn * nbm), 1, 1), st, p);
}

// aux stream of the engine (ozaki.cu)
int oz_aux_streams(Ctx* ctx);

// C(N1 x N2) = alpha * X(m x N1)^T Y(m x N2) + beta * C [+ gram_out = Y^T Y, tiles touching the upper triangle]
// SD digits of both operands are produced; the main tiles run the digit pairs of the first sp_main anti-diagonals.
template <int SD, typename T>
static int oz2_tn(Ctx* ctx, int64_t m, int64_t N1, int64_t N2, double alpha, const T* X, int64_t ldx, const T* Y, int64_t ldy, double beta, T* C, int64_t ldc,
                  double* x_sumsq_out, T* gram_out, int64_t ldg, int sp_main) {
    using Cfg = OzCfg<SD>;
    RLB_CHECK((oz2_configure<SD, T, double, true>(ctx)));
    RLB_CHECK(oz_aux_streams(ctx));
    const int64_t L = oz2_chunk_rows(m);
    const int64_t nchunks = (m + L - 1) / L;
    const int nkb = (int)(L / OZ_KB);
    const int nbm = (int)((N2 + OZ_BM - 1) / OZ_BM), nbn = (int)((N1 + OZ_BN - 1) / OZ_BN);
    const int nbg = gram_out ? (int)((N2 + OZ_BN - 1) / OZ_BN) : 0;
    int gram_tiles = 0;
    for (int y = 0; y < nbg; ++y) for (int x = 0; x < nbm; ++x) gram_tiles += (y * OZ_BN <= x * OZ_BM + OZ_BM - 1) ? 1 : 0;
    // chunks per launch: fill whole waves of CTAs, bounded by ~1.2 GB of M-side digits per buffer
    const int64_t per_chunk = (int64_t)nbm * nbn + gram_tiles;
    const int64_t ys = (int64_t)nbm * nkb * SD * OZ_TILE_A;             // bytes of Y digits per chunk
    int64_t gmax = std::max<int64_t>(1, std::min<int64_t>(nchunks, std::min<int64_t>(((int64_t)1200 << 20) / ys, 64)));
    int G = 1;
    double best = -1.0;
    for (int64_t gq = 1; gq <= gmax; ++gq) {
        const int64_t ctas = gq * per_chunk, waves = (ctas + ctx->num_sms - 1) / ctx->num_sms;
        const double eff = (double)ctas / (double)(waves * ctx->num_sms);
        if (eff >= best - 1e-9) { best = eff; G = (int)gq; }
    }
    const int nbuf = nchunks > G ? 2 : 1;
    ArenaScope as(ctx);
    const int64_t total = N1 * N2;
    double* part = as.take<double>((size_t)G * total); if (!part) return RLB200_ERR_ALLOC;
    double* part_g = nullptr;
    if (gram_out) {
        part_g = as.take<double>((size_t)G * N2 * N2); if (!part_g) return RLB200_ERR_ALLOC;
        RLB_CUDA_OK(ctx, cudaMemsetAsync(part_g, 0, sizeof(double) * (size_t)G * N2 * N2, ctx->stream));
    }
    cudaStream_t main = ctx->stream, aux = ctx->aux_stream;
    const int* Ex = nullptr;
    double* ssx = nullptr;
    if (ctx->oz_const_ptr == (const void*)X) {
        if (!oz2_cache_hit(ctx->oz2_col, X, m, N1, ldx, L, (int)sizeof(T))) RLB_CHECK(oz2_fill_const_cache<T>(ctx, X, m, N1, ldx, main));
        Ex = ctx->oz2_col.E; ssx = ctx->oz2_col.ss;
    } else {
        int* e = as.take<int>((size_t)nchunks * N1); if (!e) return RLB200_ERR_ALLOC;
        if (x_sumsq_out) { ssx = as.take<double>((size_t)nchunks * N1); if (!ssx) return RLB200_ERR_ALLOC; }
        LaunchScope ls(ctx, RLB200_TIMER_I8_SLICE, 1);
        oz_colexp_kernel<T><<<(unsigned)((N1 * nchunks + 7) / 8), 256, 0, main>>>(X, ldx, m, (int)N1, L, (int)nchunks, OZ_RAW_P, e, ssx);
        RLB_CUDA_OK(ctx, cudaGetLastError());
        Ex = e;
    }
    int* Ey = as.take<int>((size_t)nchunks * N2); if (!Ey) return RLB200_ERR_ALLOC;
    int8_t* yt[2];
    for (int b = 0; b < nbuf; ++b) { yt[b] = as.take<int8_t>((size_t)G * ys); if (!yt[b]) return RLB200_ERR_ALLOC; }
    // digits of Y: sliced on the second stream, one launch group ahead of the tensor-core kernel
    RLB_CUDA_OK(ctx, cudaEventRecord(ctx->aux_ev[0], main));
    RLB_CUDA_OK(ctx, cudaStreamWaitEvent(aux, ctx->aux_ev[0], 0));
    {
        LaunchScope ls(ctx, RLB200_TIMER_I8_SLICE, 1, aux);
        oz_colexp_kernel<T><<<(unsigned)((N2 * nchunks + 7) / 8), 256, 0, aux>>>(Y, ldy, m, (int)N2, L, (int)nchunks, Cfg::P, Ey, nullptr);
        RLB_CUDA_OK(ctx, cudaGetLastError());
    }
    int64_t it = 0;
    for (int64_t c0 = 0; c0 < nchunks; c0 += G, ++it) {
        const int g = (int)std::min<int64_t>(G, nchunks - c0);
        const int b = (int)(it % nbuf);
        if (it >= nbuf) RLB_CUDA_OK(ctx, cudaStreamWaitEvent(aux, ctx->aux_ev[3 + b], 0));
        {
            LaunchScope lsl(ctx, RLB200_TIMER_I8_SLICE, g, aux);
            for (int q = 0; q < g; ++q) {
                const int64_t r0 = (c0 + q) * L, klen = std::min(L, m - r0);
                oz_slice_tn_kernel<SD, OZ_BM, T><<<dim3((nkb + 3) / 4, nbm), 128, 0, aux>>>(Y + r0, ldy, klen, (int)N2, nkb, Ey + (c0 + q) * N2, yt[b] + q * ys);
            }
            RLB_CUDA_OK(ctx, cudaGetLastError());
        }
        RLB_CUDA_OK(ctx, cudaEventRecord(ctx->aux_ev[1 + b], aux));
        RLB_CUDA_OK(ctx, cudaStreamWaitEvent(main, ctx->aux_ev[1 + b], 0));
        {
            Oz2Params p{};
            p.m_tiles = yt[b]; p.m_group_stride = ys; p.nkb_stride = nkb;
            p.Em = Ey + c0 * N2; p.em_stride = N2; p.rows_m = (int)N2;
            p.X = X + c0 * L; p.ldx = ldx; p.rows_n = N1; p.kdim = m - c0 * L; p.L = L;
            p.En = Ex + c0 * N1; p.en_stride = N1;
            p.nkb = nkb; p.nbm = nbm; p.nbn_main = nbn; p.sp_main = sp_main;
            p.out = part; p.ldo = N1; p.out_group_stride = total; p.alpha = 1.0; p.beta = c0 > 0 ? 1.0 : 0.0;
            p.gram_out = part_g; p.gram_group_stride = N2 * N2;
            LaunchScope ls(ctx, RLB200_TIMER_I8_MMA_TN);
            RLB_CHECK((oz2_launch<SD, T, double, true>(ctx, dim3((unsigned)(nbm * (nbn + nbg)), 1, (unsigned)g), main, p)));
        }
        RLB_CUDA_OK(ctx, cudaEventRecord(ctx->aux_ev[3 + b], main));
    }
    LaunchScope ls(ctx, RLB200_TIMER_I8_MMA_TN);
    oz_reduce_kernel<T><<<(unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->num_sms * 8), 256, 0, main>>>(
        part, (int)std::min<int64_t>(G, nchunks), total, (int)N1, alpha, beta, C, ldc);
    RLB_CUDA_OK(ctx, cudaGetLastError());
    if (gram_out) {
        ctx->launches += 1;
        oz_reduce_kernel<T><<<(unsigned)std::min<int64_t>((N2 * N2 + 255) / 256, (int64_t)ctx->num_sms * 8), 256, 0, main>>>(
            part_g, (int)std::min<int64_t>(G, nchunks), N2 * N2, (int)N2, 1.0, 0.0, gram_out, ldg);
        RLB_CUDA_OK(ctx, cudaGetLastError());
    }
    if (x_sumsq_out) {
        ctx->launches += 1;
        oz_sum_kernel<<<1, 1024, 0, main>>>(ssx, nchunks * N1, x_sumsq_out);
        RLB_CUDA_OK(ctx, cudaGetLastError());
    }
    return 0;
}

static int oz2_digits(Ctx* ctx, size_t elem) {
    if (ctx->i8_digits > 0) return ctx->i8_digits;
    return elem == 8 ? 6 : 4;
}

// the raw tiles travel as 16-byte aligned bulk copies: base pointer and leading dimension (in bytes) must be multiples of 16
static bool oz2_aligned(const void* A, int64_t ld_bytes) { return (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (ld_bytes & 15) == 0; }
bool ozaki2_nn_ok(Ctx* ctx, int64_t m, int64_t N, int64_t K, const void* A, int64_t lda_bytes, const void* C) {
    // in place (C == A): the CTAs of a row tile synchronise as a cluster of ceil(N / 128) <= 8 before writing; the output must then cover
    // exactly the columns that were read (N == K: U = Y M) so that no other row tile's input is touched
    const bool inplace = A == C;
    const int64_t nbm_ = (N + OZ_BM - 1) / OZ_BM;
    if (inplace && !(N == K && (nbm_ == 1 || nbm_ == 2 || nbm_ == 4))) return false;
    return ctx->i8_fused && m > 0 && N >= 96 && K > 0 && K <= OZ_KMAX && oz2_aligned(A, lda_bytes);
}
bool ozaki2_tn_ok(Ctx* ctx, int64_t m, int64_t N1, int64_t N2, const void* X, int64_t ldx_bytes) {
    return ctx->i8_fused && m > 0 && N1 > 0 && N2 >= 96 && oz2_aligned(X, ldx_bytes);
}

template <typename T>
int ozaki2_gemm_nn(Ctx* ctx, int64_t m, int64_t N, int64_t K, double alpha, const T* A, int64_t lda, const T* B, int64_t ldb, double beta, T* C, int64_t ldc) {
    RLB_REQUIRE(ctx, (ozaki2_nn_ok(ctx, m, N, K, A, lda * (int64_t)sizeof(T), C)) && N < (1 << 20));
    RLB_REQUIRE(ctx, (const void*)A != (const void*)C || lda == ldc);
    switch (oz2_digits(ctx, sizeof(T))) {
        case 3: return oz2_nn<3, T>(ctx, m, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
        case 4: return oz2_nn<4, T>(ctx, m, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
        case 5: return oz2_nn<5, T>(ctx, m, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
        case 6: return oz2_nn<6, T>(ctx, m, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
        default: return oz2_nn<7, T>(ctx, m, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    }
}
// gram_out != nullptr: one digit more than the default is produced (fp64: 7, 54 bits) and the Gram tiles run all of its digit pairs;
// the X^T Y tiles run all of them too when full_pairs is set (28 pairs, fp64-level: the power-iteration sketch Omega = (A^T Y) R^-1
// amplifies their error by cond(R)), else the default number of anti-diagonals (21 pairs: B^T = A^T Q, whose error only reaches sigma)
template <typename T>
int ozaki2_gemm_tn(Ctx* ctx, int64_t m, int64_t N1, int64_t N2, double alpha, const T* X, int64_t ldx, const T* Y, int64_t ldy, double beta, T* C, int64_t ldc,
                   double* x_sumsq_out, T* gram_out, int64_t ldg, bool full_pairs) {
    RLB_REQUIRE(ctx, (ozaki2_tn_ok(ctx, m, N1, N2, X, ldx * (int64_t)sizeof(T))) && N1 < (1 << 20) && N2 < (1 << 20));
    const int s0 = oz2_digits(ctx, sizeof(T));
    const bool user = ctx->i8_digits > 0;
    const int sd = (gram_out && !user && sizeof(T) == 8) ? 7 : s0;
    const int sp = (gram_out && !user && sizeof(T) == 8 && !full_pairs) ? 6 : sd;
    switch (sd) {
        case 3: return oz2_tn<3, T>(ctx, m, N1, N2, alpha, X, ldx, Y, ldy, beta, C, ldc, x_sumsq_out, gram_out, ldg, sp);
        case 4: return oz2_tn<4, T>(ctx, m, N1, N2, alpha, X, ldx, Y, ldy, beta, C, ldc, x_sumsq_out, gram_out, ldg, sp);
        case 5: return oz2_tn<5, T>(ctx, m, N1, N2, alpha, X, ldx, Y, ldy, beta, C, ldc, x_sumsq_out, gram_out, ldg, sp);
        case 6: return oz2_tn<6, T>(ctx, m, N1, N2, alpha, X, ldx, Y, ldy, beta, C, ldc, x_sumsq_out, gram_out, ldg, sp);
        default: return oz2_tn<7, T>(ctx, m, N1, N2, alpha, X, ldx, Y, ldy, beta, C, ldc, x_sumsq_out, gram_out, ldg, sp);
    }
}
template int ozaki2_gemm_nn<double>(Ctx*, int64_t, int64_t, int64_t, double, const double*, int64_t, const double*, int64_t, double, double*, int64_t);
template int ozaki2_gemm_nn<float>(Ctx*, int64_t, int64_t, int64_t, double, const float*, int64_t, const float*, int64_t, double, float*, int64_t);
template int ozaki2_gemm_tn<double>(Ctx*, int64_t, int64_t, int64_t, double, const double*, int64_t, const double*, int64_t, double, double*, int64_t, double*,
                                    double*, int64_t, bool);
template int ozaki2_gemm_tn<float>(Ctx*, int64_t, int64_t, int64_t, double, const float*, int64_t, const float*, int64_t, double, float*, int64_t, double*,
                                   float*, int64_t, bool);

}  // namespace rlb
