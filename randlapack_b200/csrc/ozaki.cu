// Tall-skinny GEMMs of the path on the 5th-generation tensor cores: tcgen05.mma.kind::i8 on error-free int8 digit slices
// (Ozaki scheme), for fp64 and fp32 storage.
//
// sm_100a has no fp64 kind of tcgen05.mma; its fp64 pipe (DMMA, gemm.cu) peaks at 37 TFLOP/s measured, its int8 tensor pipe at
// 4.5 Pop/s nominal.  An operand value x with a power-of-two scale 2^E bounding its group (|x| < 2^E) is written EXACTLY as
// fixed-point balanced base-256 digits
//     rn(x * 2^(P-E)) = F = sum_t d_t * 256^(S-1-t),   d_t in [-128, 127] (int8),  P = 8S - 2,
// so a product over K terms becomes S(S+1)/2 int8 GEMMs (digit pairs s + t <= S - 1) whose int32 accumulation in TMEM is exact
// (|d d'| <= 2^14, <= S products per anti-diagonal and k, K <= 2^14 per accumulation group), recombined in fp64 in the epilogue:
//     C_ij = 2^(EA_i + EB_j - 12) * sum_d 256^(-d) * acc_d(i, j).
// S = 6 (default for fp64) keeps 46 bits below each group's largest magnitude at 21 int8 MMAs per fp64 MMA-equivalent
// (4.5 P / 21 = 214 TF nominal ceiling); S = 7 keeps 54 bits (every mantissa bit of entries within 2x of the group maximum) at 28;
// S = 4 holds an fp32 mantissa with 6 guard bits at 10.
//
//   NN  (rl_rs.hh:153, rl_rf.hh:123):  C(m x N) = A(m x K) B(K x N)        scales: per row of A, per column of B
//   TN  (rl_rs.hh:165, rl_qb.hh:218):  C(N1 x N2) = X(m x N1)^T Y(m x N2)   scales: per column and per chunk of L rows (int32 range),
//                                                                            fp64 accumulation across chunks in a fixed order
//
// Pipeline per product: (1) exponent pre-pass (biased-exponent maxima); (2) slicer kernels write the digits to HBM pre-tiled in the
// tensor core's no-swizzle K-major core-matrix order (8 rows x 16 bytes), one contiguous block per pipeline stage, so the GEMM
// kernel needs no tensor maps; (3) ozaki_mma_kernel: a producer thread streams stages with cp.async.bulk + mbarrier, an issuer thread
// drives the tensor core.  The S digit tiles of the second operand are contiguous in a stage, i.e. they form ONE stacked K-major
// tile of 64*S rows, so digit s of the first operand is multiplied with digits 0..S-1-s of the second in ceil((S-s)/4)
// instructions of N <= 256 whose output columns are exactly the accumulators of anti-diagonals s..S-1 (S*64 TMEM columns):
// 8 instructions per K step for S = 6 instead of 21, and the first-operand tile is read from shared memory 8 times, not 21.
// Eight warps run the fp64 epilogue from tcgen05.ld.
#include "drivers.cuh"
#include "oz_common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace rlb {

struct OzGram {            // fused Gram output of the TN product (see the kernel)
    int nb_main;
    double* out2;
    int64_t out2_group_stride;
};

// grid: (second-operand row blocks, first-operand row blocks, groups) — the CTAs that share the (larger) first-operand tiles are
// adjacent in launch order, so those tiles are fetched from HBM once and hit L2 for the other N tiles.  Group g (TN: an
// accumulation chunk; NN: always 0) uses the digit tiles a_tiles + g * a_group_stride (tile-row block blockIdx.y) and
// b_tiles + g * b_group_stride (block blockIdx.x), nkb K blocks each.
// Output: out[g * out_group_stride + i + j * ldo] = alpha * 2^(Ea[g*ea_stride + i] + Eb[g*eb_stride + j] - 12) * sum + beta * out
// for i < rows_a, j < rows_b.
// Launched with a thread-block cluster of cs CTAs along x (cs = 1, 2 or 4): the cs CTAs share the first-operand tiles, so each loads
// 1/cs of every first-operand stage and multicasts it to the whole cluster (L2 -> SM traffic per CTA and K step drops from
// S * 6 KB to S * (4/cs + 2) KB); a stage slot is free again when the tensor cores of ALL cs CTAs have consumed it (multicast commit).
template <int S, typename TO, bool MN>
__global__ void __launch_bounds__(256, 1)
ozaki_mma_kernel(const int8_t* __restrict__ a_tiles, int64_t a_group_stride, const int8_t* __restrict__ b_tiles, int64_t b_group_stride, int nkb,
                 const int* __restrict__ Ea, int64_t ea_stride, const int* __restrict__ Eb, int64_t eb_stride, int64_t rows_a, int rows_b,
                 TO* __restrict__ out, int64_t ldo, int64_t out_group_stride, double alpha, double beta, long long* __restrict__ dbg, int flags,
                 OzGram gp) {
    using Cfg = OzCfg<S>;
    // Fused Gram matrix (TN, MN-major only): tile rows blockIdx.y >= gp.nb_main compute Y^T Y next to X^T Y - their first operand is
    // assembled from two 64-column tiles of the SECOND operand's digits, their output goes to gp.out2 (rows_b x rows_b per group).
    const bool gram = gp.out2 != nullptr && (int)blockIdx.y >= gp.nb_main;
    const int by = gram ? (int)blockIdx.y - gp.nb_main : (int)blockIdx.y;
    if (gram) {
        if (by * OZ_BM > (int)blockIdx.x * OZ_BN + OZ_BN - 1) return;      // upper tiles only
        rows_a = rows_b; Ea = Eb; ea_stride = eb_stride;
        out = reinterpret_cast<TO*>(gp.out2); ldo = rows_b; out_group_stride = gp.out2_group_stride;
    }
    // flags & 1: the second operand is an upper-triangular K x N matrix (tile column block x only has non-zeros in K < 64 (x + 1)):
    //            the K loop stops there.  flags & 2: only tiles that touch the upper triangle of the output are computed (Gram).
    if ((flags & 2) && by * OZ_BM > (int)blockIdx.x * OZ_BN + OZ_BN - 1) return;
    const int nkb_stride = nkb;                     // K blocks per tile row in memory
    if (flags & 1) nkb = min(nkb, ((int)blockIdx.x + 1) * (OZ_BN / OZ_KB));
    constexpr int STAGES = Cfg::STAGES;
    long long t_start = 0, t_ready = 0, t_first = 0, t_acc = 0;
    if (dbg) t_start = clock64();
    extern __shared__ __align__(1024) unsigned char oz_smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[STAGES], bar_empty[STAGES], bar_acc;
    __shared__ uint32_t tmem_base_sh;
    __shared__ int s_eb20[OZ_BN], s_ebmin, s_ebmax;      // column exponents << 20 and their range over the tile
    const int tid = threadIdx.x, warp = tid >> 5;
    const int g = blockIdx.z;
    uint32_t cs, crank;
    asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(cs));
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const uint16_t cmask = (uint16_t)((1u << cs) - 1u);
    const int8_t* ga = a_tiles + g * a_group_stride + (int64_t)by * nkb_stride * (S * OZ_TILE_A);
    const int8_t* gb = b_tiles + g * b_group_stride + (int64_t)blockIdx.x * nkb_stride * (S * OZ_TILE_B);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(oz_smem(&bar_full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(oz_smem(&bar_empty[s])), "r"(cs));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(oz_smem(&bar_acc)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem(&tmem_base_sh)), "n"(Cfg::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else if (warp == 2) {
        // column exponents of this tile (out-of-range columns mirror an in-range one so that they do not widen the range)
        const int l = tid & 31;
        const int j0 = blockIdx.x * OZ_BN + l, j1 = j0 + 32;
        const int e0 = Eb[g * eb_stride + min(j0, rows_b - 1)], e1 = Eb[g * eb_stride + min(j1, rows_b - 1)];
        s_eb20[l] = e0 << 20; s_eb20[l + 32] = e1 << 20;
        const int mn = __reduce_min_sync(0xffffffffu, min(e0, e1)), mx = __reduce_max_sync(0xffffffffu, max(e0, e1));
        if (l == 0) { s_ebmin = mn; s_ebmax = mx; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (cs > 1) oz_cluster_sync();      // every CTA's barriers are initialised before any peer multicasts into them
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_sh;
    const uint32_t sbase = oz_smem(oz_smem_raw);
    if (dbg) t_ready = clock64();

    if (tid == 0) {
        // ---- producer: one bulk copy per operand per stage
        for (int kb = 0; kb < nkb; ++kb) {
            const int slot = kb % STAGES;
            if (kb >= STAGES) oz_mbar_wait(oz_smem(&bar_empty[slot]), (uint32_t)(((kb / STAGES) - 1) & 1));
            const uint32_t bar = oz_smem(&bar_full[slot]);
            const uint32_t dst = sbase + slot * Cfg::STAGE_BYTES;
            if (gram) {
                // first operand = columns [128 by, 128 by + 128) of Y = second-operand tiles 2 by and 2 by + 1 (the MN-major layout of a
                // 128-column tile is two 64-column tiles back to back, per digit)
                const bool has1 = 2 * by + 1 < (int)gridDim.x;
                const int8_t* y0 = b_tiles + g * b_group_stride + ((int64_t)(2 * by) * nkb_stride + kb) * (S * OZ_TILE_B);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                             ::"r"(bar), "r"((uint32_t)(S * OZ_TILE_B * (has1 ? 3 : 2))) : "memory");
                for (int t = 0; t < S; ++t) {
                    oz_bulk_load(dst + t * OZ_TILE_A, y0 + t * OZ_TILE_B, OZ_TILE_B, bar);
                    if (has1) oz_bulk_load(dst + t * OZ_TILE_A + OZ_TILE_B, y0 + (int64_t)nkb_stride * (S * OZ_TILE_B) + t * OZ_TILE_B, OZ_TILE_B, bar);
                }
                oz_bulk_load(dst + S * OZ_TILE_A, gb + (int64_t)kb * (S * OZ_TILE_B), S * OZ_TILE_B, bar);
                continue;
            }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)Cfg::STAGE_BYTES) : "memory");
            if (cs == 1) {
                oz_bulk_load(dst, ga + (int64_t)kb * (S * OZ_TILE_A), S * OZ_TILE_A, bar);
            } else {
                const uint32_t piece = (uint32_t)(S * OZ_TILE_A) / cs;
                oz_bulk_load_mc(dst + crank * piece, ga + (int64_t)kb * (S * OZ_TILE_A) + crank * piece, piece, bar, cmask);
            }
            oz_bulk_load(dst + S * OZ_TILE_A, gb + (int64_t)kb * (S * OZ_TILE_B), S * OZ_TILE_B, bar);
        }
    } else if (warp == 1) {
        // ---- issuer.  s32 accumulate, signed int8 A and B, both K-major, M = 128 (cute/arch/mma_sm100_desc.hpp InstrDescriptor).
        // The whole warp runs the loop (warp-uniform operands: the descriptors stay in uniform registers and there is no per-thread
        // serialisation loop around tcgen05.mma); one elected lane issues the MMAs and the commits (tools/peaks_i8.cu: <= 45 cycles per
        // instruction instead of ~110 from a single divergent thread).
        constexpr uint32_t IDESC0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BM >> 4) << 24) | (MN ? ((1u << 15) | (1u << 16)) : 0u);
        constexpr uint64_t HI = ((uint64_t)((MN ? 512 : 256) >> 4) | ((uint64_t)1 << 14)) << 32;
        uint32_t elected = 0;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(elected));
        const uint32_t base_lo0 = sbase >> 4;
        constexpr uint32_t LBO = (uint32_t)(128 >> 4) << 16;      // (start-address field = bits 4..17 of the address: a CTA's window inside a cluster does not start at 0)
        for (int kb = 0; kb < nkb;) {
            const int slot = kb % STAGES, slot2 = (kb + 1) % STAGES;
            oz_mbar_wait(oz_smem(&bar_full[slot]), (uint32_t)((kb / STAGES) & 1));
            // up to two stages per issue round (the second only if it has landed already): the barrier wait between two rounds is serial time in
            // which the tensor pipe only has what is queued
            bool two = kb > 0 && kb + 1 < nkb && oz_mbar_test(oz_smem(&bar_full[slot2]), (uint32_t)(((kb + 1) / STAGES) & 1));
            two = __all_sync(0xffffffffu, two);
            if (dbg && kb == 0 && elected) dbg[((int64_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + 2] = clock64();
            asm volatile("tcgen05.fence::after_thread_sync;");
            if (elected) {
#pragma unroll 1
                for (int r = 0; r < (two ? 2 : 1); ++r) {
                    const int sl = r ? slot2 : slot;
                    const uint32_t lo = base_lo0 + (uint32_t)sl * (uint32_t)(Cfg::STAGE_BYTES >> 4);
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        const uint64_t da = HI | (uint64_t)(((lo + (uint32_t)((s * OZ_TILE_A) >> 4)) & 0x3FFFu) | LBO);
#pragma unroll
                        for (int t0 = 0; t0 < S - s; t0 += 4) {
                            const int nt = (S - s - t0) < 4 ? (S - s - t0) : 4;      // digit tiles of the second operand in this instruction
                            const uint32_t idesc = IDESC0 | ((uint32_t)((nt * OZ_BN) >> 3) << 17);
                            const uint64_t db = HI | (uint64_t)(((lo + (uint32_t)((S * OZ_TILE_A + t0 * OZ_TILE_B) >> 4)) & 0x3FFFu) | LBO);
                            oz_mma_i8(tmem + (uint32_t)((s + t0) * OZ_BN), da, db, idesc, (kb + r > 0 || s > 0) ? 1u : 0u);
                        }
                    }
                    if (cs == 1)
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
    }
    __syncwarp();
    // ---- epilogue: TMEM lane = tile row.  Warps w and w + 4 share lane quarter w % 4 (a warp may only touch its own quarter) and
    // take 32 of the 64 columns each.
    oz_mbar_wait_sleep(oz_smem(&bar_acc), 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (dbg) t_acc = clock64();
    const int quarter = warp & 3, chalf = warp >> 2;
    const int64_t i = (int64_t)by * OZ_BM + quarter * 32 + (tid & 31);
    const int ea = (i < rows_a) ? Ea[g * ea_stride + i] : 0;
    constexpr int ESHIFT = (2 * Cfg::P - 16 * (S - 1)) + 16;
    // fast scaling: when 2^(ea + eb - ESHIFT) is a normal double for every column of the tile, its high word is one integer add
    // an Inf/NaN entry makes its group's exponent 0x7ff - 1022: everything that depends on that group is written as NaN (the BLAS
    // path of the reference propagates it the same way)
    constexpr int NONFINITE_E = 0x7ff - 1022;
    const bool fast = (ea + s_ebmin - ESHIFT >= -1022) && (ea + s_ebmax - ESHIFT <= 1023) && ea != NONFINITE_E && s_ebmax != NONFINITE_E;
    const int ea_hi = (ea - ESHIFT + 1023) << 20;
    const bool simple = fast && alpha == 1.0 && beta == 0.0 && (int)(blockIdx.x + 1) * OZ_BN <= rows_b;
    TO* og = out + g * out_group_stride;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    for (int c0 = chalf * 32; c0 < chalf * 32 + 32; c0 += 8) {
        // all S diagonals of 8 columns in flight, one wait
        uint32_t r[S][8];
#pragma unroll
        for (int d = 0; d < S; ++d) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r[d][0]), "=r"(r[d][1]), "=r"(r[d][2]), "=r"(r[d][3]), "=r"(r[d][4]), "=r"(r[d][5]), "=r"(r[d][6]), "=r"(r[d][7])
                         : "r"(lane_addr + (uint32_t)(d * OZ_BN + c0)));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        auto combine = [&](int j) -> double { return oz_combine<S>(&r[0][j]); };      // 2^16 * sum_d acc_d 256^-d
        if (i < rows_a) {
            if (simple) {
                TO* p = og + i + (int64_t)(blockIdx.x * OZ_BN + c0) * ldo;
#pragma unroll
                for (int j = 0; j < 8; ++j) p[(int64_t)j * ldo] = (TO)(combine(j) * __hiloint2double(ea_hi + s_eb20[c0 + j], 0));
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int jj = blockIdx.x * OZ_BN + c0 + j;
                    if (jj < rows_b) {
                        const double v = combine(j);
                        double val;
                        if (fast) {
                            val = v * __hiloint2double(ea_hi + s_eb20[c0 + j], 0);
                        } else if (ea == NONFINITE_E || (s_eb20[c0 + j] >> 20) == NONFINITE_E) {
                            val = __longlong_as_double(0x7ff8000000000000ll);
                        } else {
                            // v * 2^(ea + eb - ESHIFT), split in two exact power-of-two factors so that neither leaves the normal range early
                            const int e = ea + (s_eb20[c0 + j] >> 20) - ESHIFT;
                            const int e1 = e / 2, e2 = e - e1;
                            val = (v * oz_pow2(max(-1022, min(1023, e1)))) * oz_pow2(max(-1022, min(1023, e2)));
                        }
                        if (alpha != 1.0) val *= alpha;
                        TO* p = og + i + (int64_t)jj * ldo;
                        if (beta != 0.0) val += beta * (double)(*p);
                        *p = (TO)val;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(Cfg::TMEM_COLS));
    if (dbg && tid == 0) {
        long long* d = dbg + ((int64_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8;
        d[0] = t_start; d[1] = t_ready; d[3] = t_acc; d[4] = clock64();
        uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        d[5] = smid; d[6] = t_first;
    }
    if (cs > 1) oz_cluster_sync();      // no CTA exits while a peer's commit may still arrive on its barriers
}


// cluster size along x for a grid of nxb second-operand blocks: the largest of {4, 2, 1} that divides nxb and still lets
// (almost) every SM hold a CTA (GPCs whose SM count is not a multiple of the cluster size strand SMs).  RLB200_OZ_CLUSTER overrides.
template <int S, typename TO, bool MN>
static int oz_configure(Ctx* ctx, int* cs_ok /* [5] */) {
    // function attributes are per device: one flag per device id (one process may drive several contexts)
    static bool done_dev[64] = {};
    static int ok[5] = {0, 1, 0, 0, 0};
    bool& done = done_dev[ctx->device & 63];
    if (!done) {
        const int smem = OzCfg<S>::STAGES * OzCfg<S>::STAGE_BYTES;
        RLB_CUDA_OK(ctx, cudaFuncSetAttribute(ozaki_mma_kernel<S, TO, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        for (int cs = 2; cs <= 4; cs *= 2) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(cs * ctx->num_sms, 1, 1);
            cfg.blockDim = dim3(256);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int ncl = 0;
            if (cudaOccupancyMaxActiveClusters(&ncl, ozaki_mma_kernel<S, TO, MN>, &cfg) != cudaSuccess) { cudaGetLastError(); ncl = 0; }
            ok[cs] = (ncl * cs * 100 >= ctx->num_sms * 97) ? 1 : 0;
        }
        // measured (profiles/): the kernel is bound by shared-memory bandwidth, not by L2 -> SM traffic, so multicast buys nothing and
        // cluster scheduling costs a little; clusters stay opt-in (RLB200_OZ_CLUSTER = 2 or 4)
        const char* e = getenv("RLB200_OZ_CLUSTER");
        const int force = e ? atoi(e) : 1;
        for (int cs = 2; cs <= 4; cs *= 2) ok[cs] = (ok[cs] && cs <= force) ? 1 : 0;
        done = true;
    }
    for (int i = 0; i < 5; ++i) cs_ok[i] = ok[i];
    return 0;
}
static int oz_pick_cluster(const int* cs_ok, int nxb) {
    if (cs_ok[4] && nxb % 4 == 0) return 4;
    if (cs_ok[2] && nxb % 2 == 0) return 2;
    return 1;
}
template <int S, typename TO, bool MN>
static int oz_launch_mma(Ctx* ctx, dim3 grid, int cs, cudaStream_t stream, const int8_t* a_tiles, int64_t a_group_stride, const int8_t* b_tiles,
                         int64_t b_group_stride, int nkb, const int* Ea, int64_t ea_stride, const int* Eb, int64_t eb_stride, int64_t rows_a, int rows_b,
                         TO* out, int64_t ldo, int64_t out_group_stride, double alpha, double beta, long long* dbg = nullptr, int flags = 0,
                         OzGram gp = OzGram{0, nullptr, 0}) {
    // flag-dependent control flow (K loop cut at the diagonal, upper-tile-only output, Gram tiles) makes the CTAs of a cluster run different
    // K-loop lengths or return early, which the multicast / commit protocol cannot tolerate: such launches never use clusters
    if (flags != 0 || gp.out2 != nullptr) cs = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = OzCfg<S>::STAGES * OzCfg<S>::STAGE_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    RLB_CUDA_OK(ctx, cudaLaunchKernelEx(&cfg, ozaki_mma_kernel<S, TO, MN>, a_tiles, a_group_stride, b_tiles, b_group_stride, nkb, Ea, ea_stride, Eb, eb_stride,
                                        rows_a, rows_b, out, ldo, out_group_stride, alpha, beta, dbg, flags, gp));
    return 0;
}
// RLB200_OZ_DBG=1: per-CTA cycle stamps of one launch (start, barriers/TMEM ready, first stage landed, accumulators complete, end),
// averaged and printed to stderr.  Diagnostic only.
static void oz_dbg_report(const char* what, cudaStream_t stream, long long* dbg, int64_t nctas) {
    cudaStreamSynchronize(stream);
    std::vector<long long> h((size_t)nctas * 8);
    cudaMemcpy(h.data(), dbg, h.size() * 8, cudaMemcpyDeviceToHost);
    double a = 0, b = 0, c = 0, d = 0;
    long long t0 = h[0], t1 = h[4];
    for (int64_t i = 0; i < nctas; ++i) {
        const long long* x = &h[(size_t)i * 8];
        a += double(x[1] - x[0]); b += double(x[2] - x[1]); c += double(x[3] - x[2]); d += double(x[4] - x[3]);
        t0 = std::min(t0, x[0]); t1 = std::max(t1, x[4]);
    }
    fprintf(stderr, "[oz dbg] %s: %lld CTAs; cycles/CTA: setup %.0f, first stage %.0f, mainloop %.0f, epilogue %.0f; kernel span %lld cycles (clock64 is per SM)\n",
            what, (long long)nctas, a / nctas, b / nctas, c / nctas, d / nctas, t1 - t0);
}

// ------------------------------------------------------------------------------------------------
// C(m x N) = alpha * A(m x K) * B(K x N) + beta * C, tall A (col-major)
// ------------------------------------------------------------------------------------------------
// second stream: slicing of chunk c + 1 overlaps the tensor-core kernel of chunk c (double-buffered digit tiles)
template <int S, typename T>
static int oz_configure_slicers(Ctx* ctx) {
    // the slicers run beside the tensor-core kernel (which needs the maximum shared-memory carveout): ask for the same carveout so
    // that both can be resident on one SM
    static bool done_dev[64] = {};
    bool& done = done_dev[ctx->device & 63];
    if (!done) {
        RLB_CUDA_OK(ctx, cudaFuncSetAttribute(oz_slice_rows_kernel<S, T>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        RLB_CUDA_OK(ctx, cudaFuncSetAttribute(oz_slice_cols_kernel<S, OZ_BM, T>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        RLB_CUDA_OK(ctx, cudaFuncSetAttribute(oz_slice_cols_kernel<S, OZ_BN, T>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        RLB_CUDA_OK(ctx, cudaFuncSetAttribute(oz_slice_tn_kernel<S, OZ_BM, T>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        RLB_CUDA_OK(ctx, cudaFuncSetAttribute(oz_slice_tn_kernel<S, OZ_BN, T>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        RLB_CUDA_OK(ctx, cudaFuncSetAttribute(oz_rowexp_kernel<T>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        RLB_CUDA_OK(ctx, cudaFuncSetAttribute(oz_colexp_kernel<T>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        done = true;
    }
    return 0;
}
static int oz_aux(Ctx* ctx) {
    if (!ctx->aux_stream) {
        // the slicers are the critical path of a pass: their CTAs go first whenever an SM has room (RLB200_OZ_AUX_PRIO=0: plain stream)
        int lo = 0, hi = 0;
        const char* pe = getenv("RLB200_OZ_AUX_PRIO");
        if ((pe == nullptr || atoi(pe) != 0) && cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess)
            RLB_CUDA_OK(ctx, cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, hi));
        else
            RLB_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
        for (auto& e : ctx->aux_ev) RLB_CUDA_OK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    return 0;
}
int oz_aux_streams(Ctx* ctx) { return oz_aux(ctx); }
enum { OZ_EV_FORK = 0, OZ_EV_SLICED = 1, OZ_EV_CONSUMED = 3 };

// cached exponents of the constant data matrix (OzConstScope, drivers.cuh)
static int oz_cache_reserve(Ctx* ctx, OzCacheEntry& e, size_t n_e, size_t n_s) {
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
    return 0;
}
static bool oz_cache_hit(const OzCacheEntry& e, const void* ptr, int64_t m, int64_t n, int64_t ld, int64_t L, int P, int elem) {
    return e.valid && e.ptr == ptr && e.m == m && e.n == n && e.ld == ld && e.L == L && e.P == P && e.elem == elem;
}
static void oz_cache_set(OzCacheEntry& e, const void* ptr, int64_t m, int64_t n, int64_t ld, int64_t L, int P, int elem) {
    e.ptr = ptr; e.m = m; e.n = n; e.ld = ld; e.L = L; e.P = P; e.elem = elem; e.valid = true;
}
void oz_cache_destroy(Ctx* ctx) {
    for (OzCacheEntry* e : {&ctx->oz_row, &ctx->oz_col}) {
        if (e->E) cudaFree(e->E);
        if (e->ss) cudaFree(e->ss);
        *e = OzCacheEntry();
    }
}

// RLB200_OZ_TIMELINE=1: event pairs around every slicing batch (second stream) and every tensor-core launch (main stream) of one
// product, printed relative to the first event.  Diagnostic only.
struct OzTimeline {
    bool on;
    std::vector<cudaEvent_t> ev;
    std::vector<char> tag;
    OzTimeline() : on(getenv("RLB200_OZ_TIMELINE") != nullptr) {}
    void mark(char t, cudaStream_t st) {
        if (!on) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); ev.push_back(e); tag.push_back(t);
    }
    void report(const char* what) {
        if (!on || ev.empty()) return;
        cudaDeviceSynchronize();
        fprintf(stderr, "[oz timeline] %s (ms from first event; s/S = slicing begin/end, m/M = mma begin/end):", what);
        for (size_t i = 0; i < ev.size() && i < 80; ++i) {
            float ms = 0; cudaEventElapsedTime(&ms, ev[0], ev[i]);
            fprintf(stderr, " %c%.3f", tag[i], ms);
        }
        float tot = 0; cudaEventElapsedTime(&tot, ev[0], ev.back());
        fprintf(stderr, " ... total %.3f\n", tot);
        for (auto e : ev) cudaEventDestroy(e);
    }
};

template <int S, typename T>
static int oz_nn(Ctx* ctx, int64_t m, int64_t N, int64_t K, double alpha, const T* A, int64_t lda, const T* B, int64_t ldb, double beta, T* C,
                 int64_t ldc, bool b_upper_tri) {
    using Cfg = OzCfg<S>;
    int cs_ok[5];
    RLB_CHECK((oz_configure<S, T, false>(ctx, cs_ok)));
    RLB_CHECK((oz_configure_slicers<S, T>(ctx)));
    RLB_CHECK(oz_aux(ctx));
    const int nkb = (int)((K + OZ_KB - 1) / OZ_KB);
    const int nnb = (int)((N + OZ_BN - 1) / OZ_BN);
    const int cs = oz_pick_cluster(cs_ok, nnb);
    // rows of A sliced per launch: ~256 MB of digits per buffer, whole 128-row blocks
    int64_t RC = std::max<int64_t>(OZ_BM, std::min<int64_t>(((m + OZ_BM - 1) / OZ_BM) * OZ_BM,
                                                            ((int64_t)(256 << 20) / ((int64_t)nkb * OZ_KB * S)) / OZ_BM * OZ_BM));
    if (m > RC) {
        // whole waves: row blocks per launch * nnb a multiple of the SM count
        int64_t ga = ctx->num_sms, gb = nnb;
        while (gb) { const int64_t t = ga % gb; ga = gb; gb = t; }
        const int64_t unit = ctx->num_sms / ga;
        const int64_t nrb_q = (RC / OZ_BM) / unit * unit;
        if (nrb_q > 0) RC = nrb_q * OZ_BM;
    }
    const int nbuf = m > RC ? 2 : 1;
    ArenaScope as(ctx);
    int* Eb = as.take<int>(N); if (!Eb) return RLB200_ERR_ALLOC;
    int8_t* bt = as.take<int8_t>((size_t)nnb * nkb * S * OZ_TILE_B); if (!bt) return RLB200_ERR_ALLOC;
    const bool cached = ctx->oz_const_ptr == (const void*)A;     // A does not change during the enclosing driver scope
    int* Ea[2]; int8_t* at[2];
    for (int b = 0; b < nbuf; ++b) {
        Ea[b] = nullptr;
        if (!cached) { Ea[b] = as.take<int>(RC); if (!Ea[b]) return RLB200_ERR_ALLOC; }
        at[b] = as.take<int8_t>((size_t)(RC / OZ_BM) * nkb * S * OZ_TILE_A); if (!at[b]) return RLB200_ERR_ALLOC;
    }
    cudaStream_t main = ctx->stream, aux = ctx->aux_stream;
    bool fill_cache = false;
    if (cached && !oz_cache_hit(ctx->oz_row, A, m, K, lda, 0, Cfg::P, (int)sizeof(T))) {
        RLB_CHECK(oz_cache_reserve(ctx, ctx->oz_row, (size_t)m, 0));
        fill_cache = true;
    }
    RLB_CUDA_OK(ctx, cudaEventRecord(ctx->aux_ev[OZ_EV_FORK], main));
    RLB_CUDA_OK(ctx, cudaStreamWaitEvent(aux, ctx->aux_ev[OZ_EV_FORK], 0));
    if (fill_cache) {
        // the same sweep also prepares what the A^T Y products of this driver scope will ask for: the column-chunk exponents (for the digit
        // count those products use) and ||A||_F^2 - one pass over A instead of two
        const int S_tn = ctx->i8_digits ? ctx->i8_digits : (sizeof(T) == 8 ? 7 : 4);
        const int P_tn = 8 * S_tn - 2;
        const int64_t L = std::min<int64_t>(OZ_CHUNK, ((m + OZ_KB - 1) / OZ_KB) * OZ_KB);
        const int64_t nchunks = (m + L - 1) / L, ncta = (m + 255) / 256;
        const bool both = (L % 256 == 0) && K <= 8192 && (size_t)nchunks * K >= (size_t)ncta && getenv("RLB200_OZ_NO_FUSED_EXP") == nullptr;
        LaunchScope ls(ctx, RLB200_TIMER_I8_SLICE, both ? 2 : 1, aux);
        if (both) {
            RLB_CHECK(oz_cache_reserve(ctx, ctx->oz_col, (size_t)nchunks * K, (size_t)nchunks * K));
            RLB_CUDA_OK(ctx, cudaMemsetAsync(ctx->oz_col.E, 0, sizeof(int) * (size_t)nchunks * K, aux));
            RLB_CUDA_OK(ctx, cudaMemsetAsync(ctx->oz_col.ss, 0, sizeof(double) * (size_t)nchunks * K, aux));
            oz_rowcolexp_kernel<T><<<(unsigned)ncta, 256, (size_t)K * sizeof(int), aux>>>(A, lda, m, (int)K, L, Cfg::P, ctx->oz_row.E, ctx->oz_col.E, ctx->oz_col.ss);
            oz_field_to_exp_kernel<<<(unsigned)((nchunks * K + 255) / 256), 256, 0, aux>>>(ctx->oz_col.E, nchunks * K, P_tn);
            oz_cache_set(ctx->oz_col, A, m, K, lda, L, P_tn, (int)sizeof(T));
        } else {
            oz_rowexp_kernel<T><<<(unsigned)((m + 255) / 256), 256, 0, aux>>>(A, lda, m, (int)K, Cfg::P, ctx->oz_row.E);
        }
        RLB_CUDA_OK(ctx, cudaGetLastError());
        oz_cache_set(ctx->oz_row, A, m, K, lda, 0, Cfg::P, (int)sizeof(T));
    }
    {
        LaunchScope ls(ctx, RLB200_TIMER_I8_SLICE, 2);
        oz_colexp_kernel<T><<<(unsigned)((N + 7) / 8), 256, 0, main>>>(B, ldb, K, (int)N, K, 1, Cfg::P, Eb, nullptr);
        oz_slice_cols_kernel<S, OZ_BN, T><<<dim3(nnb, (nkb + 7) / 8), 128, 0, main>>>(B, ldb, K, (int)N, nkb, Eb, bt);
        RLB_CUDA_OK(ctx, cudaGetLastError());
    }
    OzTimeline tl;
    int64_t c = 0;
    for (int64_t r0 = 0; r0 < m; r0 += RC, ++c) {
        const int64_t rows = std::min(RC, m - r0);
        const int nrb = (int)((rows + OZ_BM - 1) / OZ_BM);
        const int b = (int)(c % nbuf);
        if (c >= nbuf) RLB_CUDA_OK(ctx, cudaStreamWaitEvent(aux, ctx->aux_ev[OZ_EV_CONSUMED + b], 0));
        tl.mark('s', aux);
        const int* Ea_c = cached ? ctx->oz_row.E + r0 : Ea[b];
        {
            LaunchScope ls(ctx, RLB200_TIMER_I8_SLICE, cached ? 1 : 2, aux);
            if (!cached) oz_rowexp_kernel<T><<<(unsigned)((rows + 255) / 256), 256, 0, aux>>>(A + r0, lda, rows, (int)K, Cfg::P, Ea[b]);
            oz_slice_rows_kernel<S, T><<<dim3(nrb, (nkb + 7) / 8), OZ_BM, 0, aux>>>(A + r0, lda, rows, (int)K, nkb, Ea_c, at[b]);
            RLB_CUDA_OK(ctx, cudaGetLastError());
        }
        tl.mark('S', aux);
        RLB_CUDA_OK(ctx, cudaEventRecord(ctx->aux_ev[OZ_EV_SLICED + b], aux));
        RLB_CUDA_OK(ctx, cudaStreamWaitEvent(main, ctx->aux_ev[OZ_EV_SLICED + b], 0));
        tl.mark('m', main);
        {
            LaunchScope ls(ctx, RLB200_TIMER_I8_MMA_NN);
            long long* dbg = nullptr;
            if (c == 0 && getenv("RLB200_OZ_DBG")) cudaMalloc(&dbg, (size_t)nnb * nrb * 64);
            RLB_CHECK((oz_launch_mma<S, T, false>(ctx, dim3(nnb, nrb, 1), b_upper_tri ? 1 : cs, main, at[b], 0, bt, 0, nkb, Ea_c, 0, Eb, 0, rows, (int)N, C + r0, ldc, 0, alpha, beta, dbg, b_upper_tri ? 1 : 0)));
            if (dbg) { oz_dbg_report("NN", main, dbg, (int64_t)nnb * nrb); cudaFree(dbg); }
        }
        tl.mark('M', main);
        RLB_CUDA_OK(ctx, cudaEventRecord(ctx->aux_ev[OZ_EV_CONSUMED + b], main));
    }
    tl.report("NN");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// C(N1 x N2) = alpha * X(m x N1)^T * Y(m x N2) + beta * C, contraction over the long dimension
// ------------------------------------------------------------------------------------------------
template <int S, typename T>
static int oz_tn(Ctx* ctx, int64_t m, int64_t N1, int64_t N2, double alpha, const T* X, int64_t ldx, const T* Y, int64_t ldy, double beta, T* C,
                 int64_t ldc, double* x_sumsq_out, bool upper_only, T* gram_out, int64_t ldg) {
    using Cfg = OzCfg<S>;
    int cs_ok[5];
    static const bool kmajor = getenv("RLB200_OZ_TN_KMAJOR") != nullptr;      // diagnostics: the K-major column slicer
    if (kmajor) RLB_CHECK((oz_configure<S, double, false>(ctx, cs_ok)));
    else RLB_CHECK((oz_configure<S, double, true>(ctx, cs_ok)));
    RLB_CHECK((oz_configure_slicers<S, T>(ctx)));
    RLB_CHECK(oz_aux(ctx));
    const int64_t L = std::min<int64_t>(OZ_CHUNK, ((m + OZ_KB - 1) / OZ_KB) * OZ_KB);
    const int64_t nchunks = (m + L - 1) / L;
    const int nkb = (int)(L / OZ_KB);
    const int nb1 = (int)((N1 + OZ_BM - 1) / OZ_BM), nb2 = (int)((N2 + OZ_BN - 1) / OZ_BN);
    // fused Gram matrix Y^T Y: nb1g more tile rows (upper tiles only) fed from the second operand's digits
    const int nb1g = gram_out ? (int)((N2 + OZ_BM - 1) / OZ_BM) : 0;
    int gram_tiles = 0;
    for (int y = 0; y < nb1g; ++y) for (int x = 0; x < nb2; ++x) gram_tiles += (y * OZ_BM <= x * OZ_BN + OZ_BN - 1) ? 1 : 0;
    // chunks per launch: as many as fill two waves of CTAs; chunk slot q of every launch accumulates into partial q
    const int G = (int)std::max<int64_t>(1, std::min<int64_t>(nchunks, (2 * ctx->num_sms) / (nb1 * nb2 + gram_tiles)));
    const int nbuf = nchunks > G ? 2 : 1;
    const int cs = oz_pick_cluster(cs_ok, nb2);
    ArenaScope as(ctx);
    const int64_t total = N1 * N2;
    double* part = as.take<double>((size_t)G * total); if (!part) return RLB200_ERR_ALLOC;
    double* part_g = nullptr;
    if (gram_out) {
        RLB_REQUIRE(ctx, !kmajor);
        part_g = as.take<double>((size_t)G * N2 * N2); if (!part_g) return RLB200_ERR_ALLOC;
        RLB_CUDA_OK(ctx, cudaMemsetAsync(part_g, 0, sizeof(double) * (size_t)G * N2 * N2, ctx->stream));
    }
    const bool cached = ctx->oz_const_ptr == (const void*)X;     // X does not change during the enclosing driver scope
    int* Ex = nullptr;
    double* ssx = nullptr;
    bool fill_x = true;
    if (cached) {
        if (oz_cache_hit(ctx->oz_col, X, m, N1, ldx, L, Cfg::P, (int)sizeof(T))) fill_x = false;
        else RLB_CHECK(oz_cache_reserve(ctx, ctx->oz_col, (size_t)nchunks * N1, (size_t)nchunks * N1));
        Ex = ctx->oz_col.E; ssx = ctx->oz_col.ss;
    } else {
        Ex = as.take<int>((size_t)nchunks * N1); if (!Ex) return RLB200_ERR_ALLOC;
        if (x_sumsq_out) { ssx = as.take<double>((size_t)nchunks * N1); if (!ssx) return RLB200_ERR_ALLOC; }
    }
    const bool same_xy = (const void*)X == (const void*)Y && ldx == ldy && N1 == N2;     // Gram: one exponent pass serves both sides
    int* Ey = Ex;
    if (!same_xy) { Ey = as.take<int>((size_t)nchunks * N2); if (!Ey) return RLB200_ERR_ALLOC; }
    if (upper_only) RLB_CUDA_OK(ctx, cudaMemsetAsync(part, 0, sizeof(double) * (size_t)G * total, ctx->stream));
    const int64_t xs = (int64_t)nb1 * nkb * S * OZ_TILE_A, ys = (int64_t)nb2 * nkb * S * OZ_TILE_B;   // bytes per chunk
    int8_t* xt[2]; int8_t* yt[2];
    for (int b = 0; b < nbuf; ++b) {
        xt[b] = as.take<int8_t>((size_t)G * xs); if (!xt[b]) return RLB200_ERR_ALLOC;
        yt[b] = as.take<int8_t>((size_t)G * ys); if (!yt[b]) return RLB200_ERR_ALLOC;
    }
    cudaStream_t main = ctx->stream, aux = ctx->aux_stream;
    RLB_CUDA_OK(ctx, cudaEventRecord(ctx->aux_ev[OZ_EV_FORK], main));
    RLB_CUDA_OK(ctx, cudaStreamWaitEvent(aux, ctx->aux_ev[OZ_EV_FORK], 0));
    {
    LaunchScope ls(ctx, RLB200_TIMER_I8_SLICE, (fill_x ? 1 : 0) + (same_xy ? 0 : 1), aux);
    if (fill_x) {
        oz_colexp_kernel<T><<<(unsigned)((N1 * nchunks + 7) / 8), 256, 0, aux>>>(X, ldx, m, (int)N1, L, (int)nchunks, Cfg::P, Ex, ssx);
        if (cached) oz_cache_set(ctx->oz_col, X, m, N1, ldx, L, Cfg::P, (int)sizeof(T));
    }
    if (!same_xy) oz_colexp_kernel<T><<<(unsigned)((N2 * nchunks + 7) / 8), 256, 0, aux>>>(Y, ldy, m, (int)N2, L, (int)nchunks, Cfg::P, Ey, nullptr);
    RLB_CUDA_OK(ctx, cudaGetLastError());
    }
    OzTimeline tl;
    int64_t it = 0;
    for (int64_t c0 = 0; c0 < nchunks; c0 += G, ++it) {
        const int g = (int)std::min<int64_t>(G, nchunks - c0);
        const int b = (int)(it % nbuf);
        if (it >= nbuf) RLB_CUDA_OK(ctx, cudaStreamWaitEvent(aux, ctx->aux_ev[OZ_EV_CONSUMED + b], 0));
        tl.mark('s', aux);
        {
        LaunchScope lsl(ctx, RLB200_TIMER_I8_SLICE, 2 * g, aux);
        for (int q = 0; q < g; ++q) {
            const int64_t r0 = (c0 + q) * L, klen = std::min(L, m - r0);
            if (kmajor) {
                oz_slice_cols_kernel<S, OZ_BM, T><<<dim3(nb1, (nkb + 7) / 8), 128, 0, aux>>>(X + r0, ldx, klen, (int)N1, nkb, Ex + (c0 + q) * N1, xt[b] + q * xs);
                oz_slice_cols_kernel<S, OZ_BN, T><<<dim3(nb2, (nkb + 7) / 8), 128, 0, aux>>>(Y + r0, ldy, klen, (int)N2, nkb, Ey + (c0 + q) * N2, yt[b] + q * ys);
            } else {
                oz_slice_tn_kernel<S, OZ_BM, T><<<dim3((nkb + 3) / 4, nb1), 128, 0, aux>>>(X + r0, ldx, klen, (int)N1, nkb, Ex + (c0 + q) * N1, xt[b] + q * xs);
                oz_slice_tn_kernel<S, OZ_BN, T><<<dim3((nkb + 3) / 4, nb2), 128, 0, aux>>>(Y + r0, ldy, klen, (int)N2, nkb, Ey + (c0 + q) * N2, yt[b] + q * ys);
            }
        }
        RLB_CUDA_OK(ctx, cudaGetLastError());
        }
        tl.mark('S', aux);
        RLB_CUDA_OK(ctx, cudaEventRecord(ctx->aux_ev[OZ_EV_SLICED + b], aux));
        RLB_CUDA_OK(ctx, cudaStreamWaitEvent(main, ctx->aux_ev[OZ_EV_SLICED + b], 0));
        tl.mark('m', main);
        {
            LaunchScope ls(ctx, RLB200_TIMER_I8_MMA_TN);
            long long* dbg = nullptr;
            if (c0 == 0 && getenv("RLB200_OZ_DBG")) cudaMalloc(&dbg, (size_t)nb2 * nb1 * g * 64);
            if (kmajor)
                RLB_CHECK((oz_launch_mma<S, double, false>(ctx, dim3(nb2, nb1, g), upper_only ? 1 : cs, main, xt[b], xs, yt[b], ys, nkb, Ex + c0 * N1, N1, Ey + c0 * N2, N2, N1,
                                                           (int)N2, part, N1, total, 1.0, c0 > 0 ? 1.0 : 0.0, dbg, upper_only ? 2 : 0)));
            else
                RLB_CHECK((oz_launch_mma<S, double, true>(ctx, dim3(nb2, nb1 + nb1g, g), (gram_out || upper_only) ? 1 : cs, main, xt[b], xs, yt[b], ys, nkb, Ex + c0 * N1, N1,
                                                          Ey + c0 * N2, N2, N1, (int)N2, part, N1, total, 1.0, c0 > 0 ? 1.0 : 0.0, dbg, upper_only ? 2 : 0,
                                                          OzGram{nb1, part_g, N2 * N2})));
            if (dbg) { oz_dbg_report("TN", main, dbg, (int64_t)nb2 * nb1 * g); cudaFree(dbg); }
        }
        tl.mark('M', main);
        RLB_CUDA_OK(ctx, cudaEventRecord(ctx->aux_ev[OZ_EV_CONSUMED + b], main));
    }
    tl.report("TN");
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
    if (x_sumsq_out) {      // main has waited for the slicing events, which follow the exponent pass on the second stream
        ctx->launches += 1;
        oz_sum_kernel<<<1, 1024, 0, main>>>(ssx, nchunks * N1, x_sumsq_out);
        RLB_CUDA_OK(ctx, cudaGetLastError());
    }
    return 0;
}

static int oz_default_digits(Ctx* ctx, size_t elem) {
    if (ctx->i8_digits > 0) return ctx->i8_digits;
    return elem == 8 ? 6 : 4;
}

template <typename T>
int ozaki_gemm_nn(Ctx* ctx, int64_t m, int64_t N, int64_t K, double alpha, const T* A, int64_t lda, const T* B, int64_t ldb, double beta, T* C,
                  int64_t ldc, bool b_upper_tri) {
    RLB_REQUIRE(ctx, m >= 0 && N >= 0 && K >= 0 && N < (1 << 20));
    if (m == 0 || N == 0) return 0;
    if (K == 0 || K > OZ_KMAX) return gemm_nn<T>(ctx, m, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    switch (oz_default_digits(ctx, sizeof(T))) {
        case 3: return oz_nn<3, T>(ctx, m, N, K, alpha, A, lda, B, ldb, beta, C, ldc, b_upper_tri);
        case 4: return oz_nn<4, T>(ctx, m, N, K, alpha, A, lda, B, ldb, beta, C, ldc, b_upper_tri);
        case 5: return oz_nn<5, T>(ctx, m, N, K, alpha, A, lda, B, ldb, beta, C, ldc, b_upper_tri);
        case 6: return oz_nn<6, T>(ctx, m, N, K, alpha, A, lda, B, ldb, beta, C, ldc, b_upper_tri);
        default: return oz_nn<7, T>(ctx, m, N, K, alpha, A, lda, B, ldb, beta, C, ldc, b_upper_tri);
    }
}
template <typename T>
int ozaki_gemm_tn(Ctx* ctx, int64_t m, int64_t N1, int64_t N2, double alpha, const T* X, int64_t ldx, const T* Y, int64_t ldy, double beta, T* C,
                  int64_t ldc, double* x_sumsq_out, bool upper_only, T* gram_out, int64_t ldg) {
    RLB_REQUIRE(ctx, m >= 0 && N1 >= 0 && N2 >= 0 && N1 < (1 << 20) && N2 < (1 << 20));
    if (N1 == 0 || N2 == 0) return 0;
    if (m == 0) return gemm_tn<T>(ctx, 0, N1, N2, 0.0, X, ldx, Y, ldy, beta, C, ldc, 0, x_sumsq_out);
    switch (oz_default_digits(ctx, sizeof(T))) {
        case 3: return oz_tn<3, T>(ctx, m, N1, N2, alpha, X, ldx, Y, ldy, beta, C, ldc, x_sumsq_out, upper_only, gram_out, ldg);
        case 4: return oz_tn<4, T>(ctx, m, N1, N2, alpha, X, ldx, Y, ldy, beta, C, ldc, x_sumsq_out, upper_only, gram_out, ldg);
        case 5: return oz_tn<5, T>(ctx, m, N1, N2, alpha, X, ldx, Y, ldy, beta, C, ldc, x_sumsq_out, upper_only, gram_out, ldg);
        case 6: return oz_tn<6, T>(ctx, m, N1, N2, alpha, X, ldx, Y, ldy, beta, C, ldc, x_sumsq_out, upper_only, gram_out, ldg);
        default: return oz_tn<7, T>(ctx, m, N1, N2, alpha, X, ldx, Y, ldy, beta, C, ldc, x_sumsq_out, upper_only, gram_out, ldg);
    }
}
template int ozaki_gemm_nn<double>(Ctx*, int64_t, int64_t, int64_t, double, const double*, int64_t, const double*, int64_t, double, double*, int64_t, bool);
template int ozaki_gemm_nn<float>(Ctx*, int64_t, int64_t, int64_t, double, const float*, int64_t, const float*, int64_t, double, float*, int64_t, bool);
template int ozaki_gemm_tn<double>(Ctx*, int64_t, int64_t, int64_t, double, const double*, int64_t, const double*, int64_t, double, double*, int64_t, double*, bool,
                                   double*, int64_t);
template int ozaki_gemm_tn<float>(Ctx*, int64_t, int64_t, int64_t, double, const float*, int64_t, const float*, int64_t, double, float*, int64_t, double*, bool,
                                  float*, int64_t);

}  // namespace rlb
