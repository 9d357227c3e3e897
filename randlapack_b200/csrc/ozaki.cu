// fp64 tall-skinny GEMMs on the 5th-generation tensor cores: tcgen05.mma.kind::i8 on error-free int8 slices (Ozaki scheme).
//
// sm_100a has no fp64 kind of tcgen05.mma; its fp64 pipe (DMMA, gemm.cu) peaks at 37 TFLOP/s measured, its int8 tensor pipe at
// 4.5 Pop/s nominal.  An fp64 operand x with a scale 2^E bounding its group is written EXACTLY as fixed-point digits
//     x * 2^(P-E) ~ F = sum_t d_t * 128^(S-1-t),   d_t in [-64, 64] (int8),  P = 7S - 1,
// so a product over K terms becomes S(S+1)/2 int8 GEMMs (digit pairs s + t <= S - 1) whose int32 accumulation in TMEM is exact
// (|d d'| <= 2^12, K <= 2^16 per accumulation group, <= S products per anti-diagonal), recombined in fp64 in the epilogue:
//     C_ij = 2^(EA_i + EB_j - 12) * sum_d 128^(-d) * acc_d(i, j).
// S = 7 keeps 48 bits below each group's largest magnitude (truncation ~ 128^-7 = 2e-15 relative to max|a| max|b| per term):
// the result is normwise as accurate as a DGEMM, at 28 int8 MMAs per fp64 MMA-equivalent (4.5 P / 28 = 160 TF nominal ceiling).
//
//   NN  (rl_rs.hh:153, rl_rf.hh:123):  C(m x N) = A(m x K) B(K x N)        scales: per row of A, per column of B
//   TN  (rl_rs.hh:165, rl_qb.hh:218):  C(N1 x N2) = X(m x N1)^T Y(m x N2)   scales: per column and per chunk of L rows (int32 range),
//                                                                            fp64 accumulation across chunks in a fixed order
//
// Pipeline per product: (1) max-magnitude pre-pass -> exponents; (2) slicer kernels write the digits to HBM pre-tiled in the
// tensor core's no-swizzle K-major core-matrix order (8 rows x 16 bytes), one contiguous block per pipeline stage, so the GEMM
// kernel needs no tensor maps: (3) ozaki_mma_kernel: one thread streams stages with cp.async.bulk + mbarrier and issues the
// tcgen05.mma's; accumulators (S anti-diagonals x 64 columns) live in TMEM; four warps run the fp64 epilogue from tcgen05.ld.
#include "drivers.cuh"
#include <algorithm>
#include <cmath>

namespace rlb {

constexpr int OZ_S = 7;            // digits per value
constexpr int OZ_P = 7 * OZ_S - 1; // fixed-point bits below the group scale
constexpr int OZ_KB = 32;          // K bytes (= int8 elements) per stage = one tcgen05.mma K step
constexpr int OZ_BM = 128;         // UMMA M: tile rows of the first operand
constexpr int OZ_BN = 64;          // UMMA N: tile rows of the second operand
constexpr int OZ_STAGES = 5;
constexpr int OZ_TILE_A = OZ_BM * OZ_KB;   // bytes of one digit tile of the first operand
constexpr int OZ_TILE_B = OZ_BN * OZ_KB;
constexpr int OZ_STAGE_BYTES = OZ_S * (OZ_TILE_A + OZ_TILE_B);
constexpr int OZ_TMEM_COLS = 512;          // S * 64 = 448 accumulator columns -> next power of two
constexpr int64_t OZ_CHUNK = 32768;        // rows per int32 accumulation group of the TN product (7 * 2^15 * 2^12 < 2^31)

__device__ __forceinline__ uint32_t oz_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------------
// exponents
// ------------------------------------------------------------------------------------------------
// E such that |x| < 2^E for every |x| <= mx (mx > 0); mx == 0 -> 0
__device__ __forceinline__ int oz_exp_of(double mx) {
    if (!(mx > 0.0)) return 0;
    int e;
    frexp(mx, &e);
    return e;
}

// E_row[i] for rows [r0, r0 + rows) of A (col-major, lda), K columns.  One thread per row, coalesced across rows.
__global__ void __launch_bounds__(256) oz_rowmax_kernel(const double* __restrict__ A, int64_t lda, int64_t rows, int K, int* __restrict__ E) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    double mx = 0.0;
    for (int c = 0; c < K; ++c) mx = fmax(mx, fabs(A[i + (int64_t)c * lda]));
    E[i] = oz_exp_of(mx);
}
// E[chunk * ncols + c] over rows [chunk*L, (chunk+1)*L) of column c of X (col-major).  One warp per (column, chunk).
__global__ void __launch_bounds__(256) oz_colmax_kernel(const double* __restrict__ X, int64_t ldx, int64_t rows, int ncols, int64_t L, int nchunks,
                                                        int* __restrict__ E) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= (int64_t)ncols * nchunks) return;
    const int c = (int)(w % ncols), ch = (int)(w / ncols);
    const int64_t r0 = (int64_t)ch * L, r1 = min(rows, r0 + L);
    const double* x = X + (int64_t)c * ldx;
    double mx = 0.0;
    for (int64_t r = r0 + lane; r < r1; r += 32) mx = fmax(mx, fabs(x[r]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) E[(int64_t)ch * ncols + c] = oz_exp_of(mx);
}

// ------------------------------------------------------------------------------------------------
// slicers.  Digit tiles: tile (rb, kb, t) of TR rows x 32 K-bytes at ((rb * nkb + kb) * S + t) * TR * 32, inside it the byte of
// (row r, k) sits at ((r / 8) * 2 + k / 16) * 128 + (r % 8) * 16 + k % 16  (K-major, no swizzle: SBO = 256 B, LBO = 128 B).
// ------------------------------------------------------------------------------------------------
// digits of one value.  `scale` = 2^(P - E) (exact power of two), |x * scale| <= 2^P = 2^48.
//   y = x * scale + 1.5 * 2^52   puts F = rn(x * scale) (two's complement) in the low mantissa bits of y;
//   adding BIAS = sum_t 64 * 128^t makes every 7-bit field u_t = d_t + 64 non-negative, so the digits are plain shifts and masks
//   (the top field is left unmasked: it may reach 128, i.e. d_0 = 64, which is a valid int8).
// Returned packed: digit t of this value in byte lane `lane` (0..3) of w[t], already re-centred (u - 64 as two's complement byte).
constexpr long long OZ_MAGIC_BITS = 0x4338000000000000ll;                       // bit pattern of 1.5 * 2^52
constexpr long long OZ_BIAS = 64ll * ((1ll << 49) - 1) / 127;                   // sum_{t<7} 64 * 128^t
__device__ __forceinline__ void oz_digits_packed(double x, double scale, int lane, uint32_t* w /* [S] */) {
    const double y = fma(x, scale, 6755399441055744.0);
    const unsigned long long Fp = (unsigned long long)(__double_as_longlong(y) - OZ_MAGIC_BITS + OZ_BIAS);
    const int sh8 = lane * 8;
#pragma unroll
    for (int t = 0; t < OZ_S; ++t) {
        uint32_t u = (uint32_t)(Fp >> (7 * (OZ_S - 1 - t)));
        if (t > 0) u &= 127u;
        w[t] |= ((u - 64u) & 0xFFu) << sh8;
    }
}
__device__ __forceinline__ double oz_pow2(int e) {      // 2^e for e in [-1022, 1023]
    return __longlong_as_double((long long)(e + 1023) << 52);
}
// scale = 2^(P - E), clamped to the normal range (inputs below 2^-970 of magnitude lose digits, never correctness of the bound)
__device__ __forceinline__ double oz_scale_of(int E) { return oz_pow2(max(-1022, min(1023, OZ_P - E))); }

// First operand of the NN product: tile rows = rows of A, K = columns of A, scale per row.  CTA = (row block, group of 8 K-blocks).
template <int TR>
__global__ void __launch_bounds__(TR) oz_slice_rows_kernel(const double* __restrict__ A, int64_t lda, int64_t rows, int K, int nkb,
                                                           const int* __restrict__ E, int8_t* __restrict__ out) {
    const int r = threadIdx.x;
    const int64_t rb = blockIdx.x;
    const int64_t row = rb * TR + r;
    const bool rv = row < rows;
    const double scale = rv ? oz_scale_of(E[row]) : 0.0;
    for (int kb = blockIdx.y * 8; kb < min(nkb, (int)blockIdx.y * 8 + 8); ++kb) {
        int8_t* tile0 = out + ((rb * nkb + kb) * OZ_S) * (int64_t)(TR * OZ_KB);
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
            uint32_t pk[4][OZ_S];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int t = 0; t < OZ_S; ++t) pk[q][t] = 0;
            double xv[16];
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                const int col = kb * OZ_KB + kc * 16 + kk;
                xv[kk] = (rv && col < K) ? A[row + (int64_t)col * lda] : 0.0;
            }
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) oz_digits_packed(xv[kk], scale, kk & 3, pk[kk >> 2]);
            const int off = ((r >> 3) * 2 + kc) * 128 + (r & 7) * 16;
#pragma unroll
            for (int t = 0; t < OZ_S; ++t)
                *reinterpret_cast<uint4*>(tile0 + (int64_t)t * (TR * OZ_KB) + off) = make_uint4(pk[0][t], pk[1][t], pk[2][t], pk[3][t]);
        }
    }
}

// Operands whose K runs along the contiguous direction: tile rows = columns of X, K = rows [k0, k0 + klen) of X, scale per column
// (E[c], already offset to the chunk).  CTA = (column block, group of 8 K-blocks); thread = (column, 16-element K chunk).
template <int TR>
__global__ void __launch_bounds__(128) oz_slice_cols_kernel(const double* __restrict__ X, int64_t ldx, int64_t klen, int ncols, int nkb,
                                                            const int* __restrict__ E, int8_t* __restrict__ out) {
    const int64_t cb = blockIdx.x;
    for (int kb = blockIdx.y * 8; kb < min(nkb, (int)blockIdx.y * 8 + 8); ++kb) {
        int8_t* tile0 = out + ((cb * nkb + kb) * OZ_S) * (int64_t)(TR * OZ_KB);
        for (int item = threadIdx.x; item < TR * 2; item += 128) {
            const int cl = item >> 1, kc = item & 1;
            const int64_t c = cb * TR + cl;
            const bool cv = c < ncols;
            const double scale = cv ? oz_scale_of(E[c]) : 0.0;
            const int64_t kbase = (int64_t)kb * OZ_KB + kc * 16;
            const double* x = X + (cv ? c : 0) * ldx + kbase;
            uint32_t pk[4][OZ_S];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int t = 0; t < OZ_S; ++t) pk[q][t] = 0;
            double xv[16];
            if (cv && kbase + 16 <= klen && ((reinterpret_cast<uintptr_t>(x) & 15) == 0)) {
#pragma unroll
                for (int kk = 0; kk < 16; kk += 2) { const double2 t2 = *reinterpret_cast<const double2*>(x + kk); xv[kk] = t2.x; xv[kk + 1] = t2.y; }
            } else {
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) xv[kk] = (cv && kbase + kk < klen) ? x[kk] : 0.0;
            }
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) oz_digits_packed(xv[kk], scale, kk & 3, pk[kk >> 2]);
            const int off = ((cl >> 3) * 2 + kc) * 128 + (cl & 7) * 16;
#pragma unroll
            for (int t = 0; t < OZ_S; ++t)
                *reinterpret_cast<uint4*>(tile0 + (int64_t)t * (TR * OZ_KB) + off) = make_uint4(pk[0][t], pk[1][t], pk[2][t], pk[3][t]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// the tensor-core kernel
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr) {   // K-major, SWIZZLE_NONE, LBO = 128 B, SBO = 256 B, version 1
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void oz_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void oz_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// grid: (second-operand row blocks, first-operand row blocks, groups) — the CTAs that share the (larger) first-operand tiles are
// adjacent in launch order, so those tiles are fetched from HBM once and hit L2 for the other N tiles.  Group g (TN: an accumulation chunk; NN: always 0) uses the
// digit tiles a_tiles + g * a_group_stride (tile-row block blockIdx.y) and b_tiles + g * b_group_stride (block blockIdx.x), nkb K
// blocks each.  Output: out[g * out_group_stride + i + j * ldo] = alpha * 2^(Ea[g*ea_stride + i] + Eb[g*eb_stride + j] - 12) * sum + beta * out
// for i < rows_a, j < rows_b.
__global__ void __launch_bounds__(128, 1)
ozaki_mma_kernel(const int8_t* __restrict__ a_tiles, int64_t a_group_stride, const int8_t* __restrict__ b_tiles, int64_t b_group_stride, int nkb,
                 const int* __restrict__ Ea, int64_t ea_stride, const int* __restrict__ Eb, int64_t eb_stride, int64_t rows_a, int rows_b,
                 double* __restrict__ out, int64_t ldo, int64_t out_group_stride, double alpha, double beta) {
    extern __shared__ __align__(1024) unsigned char oz_smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[OZ_STAGES], bar_empty[OZ_STAGES], bar_acc;
    __shared__ uint32_t tmem_base_sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.z;
    const int8_t* ga = a_tiles + g * a_group_stride + (int64_t)blockIdx.y * nkb * (OZ_S * OZ_TILE_A);
    const int8_t* gb = b_tiles + g * b_group_stride + (int64_t)blockIdx.x * nkb * (OZ_S * OZ_TILE_B);

    if (tid == 0) {
        for (int s = 0; s < OZ_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(oz_smem(&bar_full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(oz_smem(&bar_empty[s])));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(oz_smem(&bar_acc)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem(&tmem_base_sh)), "n"(OZ_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_sh;

    if (tid == 0) {
        const uint32_t sbase = oz_smem(oz_smem_raw);
        // s32 accumulate, signed int8 A and B, both K-major, N = 64, M = 128 (cute/arch/mma_sm100_desc.hpp InstrDescriptor)
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) | ((uint32_t)(OZ_BM >> 4) << 24);
        auto load_stage = [&](int kb) {
            const int slot = kb % OZ_STAGES;
            const uint32_t bar = oz_smem(&bar_full[slot]);
            const uint32_t dst = sbase + slot * OZ_STAGE_BYTES;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)OZ_STAGE_BYTES) : "memory");
            oz_bulk_load(dst, ga + (int64_t)kb * (OZ_S * OZ_TILE_A), OZ_S * OZ_TILE_A, bar);
            oz_bulk_load(dst + OZ_S * OZ_TILE_A, gb + (int64_t)kb * (OZ_S * OZ_TILE_B), OZ_S * OZ_TILE_B, bar);
        };
        for (int kb = 0; kb < min(nkb, OZ_STAGES); ++kb) load_stage(kb);
        for (int kb = 0; kb < nkb; ++kb) {
            const int slot = kb % OZ_STAGES;
            oz_mbar_wait(oz_smem(&bar_full[slot]), (uint32_t)((kb / OZ_STAGES) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t sa = sbase + slot * OZ_STAGE_BYTES, sb = sa + OZ_S * OZ_TILE_A;
#pragma unroll
            for (int d = 0; d < OZ_S; ++d) {
#pragma unroll
                for (int s = 0; s <= d; ++s) {
                    const uint64_t da = oz_desc(sa + s * OZ_TILE_A), db = oz_desc(sb + (d - s) * OZ_TILE_B);
                    const uint32_t acc = (kb > 0 || s > 0) ? 1u : 0u;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
                                 ::"r"(tmem + (uint32_t)(d * OZ_BN)), "l"(da), "l"(db), "r"(idesc), "r"(acc));
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(oz_smem(&bar_empty[slot])) : "memory");
            // refill the slot used one iteration ago: its MMAs were committed a whole stage of tensor work earlier
            if (kb >= 1 && kb - 1 + OZ_STAGES < nkb) {
                const int ps = (kb - 1) % OZ_STAGES;
                oz_mbar_wait(oz_smem(&bar_empty[ps]), (uint32_t)(((kb - 1) / OZ_STAGES) & 1));
                load_stage(kb - 1 + OZ_STAGES);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(oz_smem(&bar_acc)) : "memory");
    }
    __syncwarp();
    // ---- epilogue: TMEM lane = tile row; warp w reads lanes [32w, 32w + 32)
    oz_mbar_wait(oz_smem(&bar_acc), 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    const int64_t i = (int64_t)blockIdx.y * OZ_BM + tid;
    const int ea = (i < rows_a) ? Ea[g * ea_stride + i] : 0;
    double* og = out + g * out_group_stride;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < OZ_BN; c0 += 16) {
        // all S diagonals of 16 columns in flight, one wait
        uint32_t r[OZ_S][16];
#pragma unroll
        for (int d = 0; d < OZ_S; ++d) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[d][0]), "=r"(r[d][1]), "=r"(r[d][2]), "=r"(r[d][3]), "=r"(r[d][4]), "=r"(r[d][5]), "=r"(r[d][6]), "=r"(r[d][7]),
                           "=r"(r[d][8]), "=r"(r[d][9]), "=r"(r[d][10]), "=r"(r[d][11]), "=r"(r[d][12]), "=r"(r[d][13]), "=r"(r[d][14]), "=r"(r[d][15])
                         : "r"(lane_addr + (uint32_t)(d * OZ_BN + c0)));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (i < rows_a) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int jj = blockIdx.x * OZ_BN + c0 + j;
                if (jj < rows_b) {
                    double v = 0.0;
#pragma unroll
                    for (int d = OZ_S - 1; d >= 0; --d) v = fma(v, 0.0078125, (double)(int32_t)r[d][j]);
                    // v * 2^(ea + eb - 12), split in two exact power-of-two factors so that neither leaves the normal range early
                    const int e = ea + Eb[g * eb_stride + jj] - 12;
                    const int e1 = e / 2, e2 = e - e1;
                    double val = alpha * ((v * oz_pow2(max(-1022, min(1023, e1)))) * oz_pow2(max(-1022, min(1023, e2))));
                    double* p = og + i + (int64_t)jj * ldo;
                    if (beta != 0.0) val += beta * (*p);
                    *p = val;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(OZ_TMEM_COLS));
}

// C = alpha * sum_g part[g] + beta * C  (fixed order)
__global__ void __launch_bounds__(256) oz_reduce_kernel(const double* __restrict__ part, int groups, int64_t total, int n1, double alpha, double beta,
                                                        double* __restrict__ C, int64_t ldc) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int g = 0; g < groups; ++g) s += part[(int64_t)g * total + e];
        double* c = C + (e % n1) + (e / n1) * ldc;
        double v = alpha * s;
        if (beta != 0.0) v += beta * (*c);
        *c = v;
    }
}

static int oz_configure(Ctx* ctx) {
    static bool done = false;
    if (!done) {
        RLB_CUDA_OK(ctx, cudaFuncSetAttribute(ozaki_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_STAGES * OZ_STAGE_BYTES));
        done = true;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// C(m x N) = alpha * A(m x K) * B(K x N) + beta * C, tall A (col-major fp64)
// ------------------------------------------------------------------------------------------------
int ozaki_gemm_nn(Ctx* ctx, int64_t m, int64_t N, int64_t K, double alpha, const double* A, int64_t lda, const double* B, int64_t ldb, double beta,
                  double* C, int64_t ldc) {
    RLB_REQUIRE(ctx, m >= 0 && N >= 0 && K >= 0 && K <= 65536 && N < (1 << 20));
    if (m == 0 || N == 0) return 0;
    if (K == 0) return gemm_nn<double>(ctx, m, N, 0, 0.0, A, lda, B, ldb, beta, C, ldc);
    RLB_CHECK(oz_configure(ctx));
    const int nkb = (int)((K + OZ_KB - 1) / OZ_KB);
    const int nnb = (int)((N + OZ_BN - 1) / OZ_BN);
    const int64_t RC = 32768;                       // rows of A sliced per launch
    ArenaScope as(ctx);
    int* Eb = as.take<int>(N); if (!Eb) return RLB200_ERR_ALLOC;
    int8_t* bt = as.take<int8_t>((size_t)nnb * nkb * OZ_S * OZ_TILE_B); if (!bt) return RLB200_ERR_ALLOC;
    int* Ea = as.take<int>(RC); if (!Ea) return RLB200_ERR_ALLOC;
    int8_t* at = as.take<int8_t>((size_t)(RC / OZ_BM) * nkb * OZ_S * OZ_TILE_A); if (!at) return RLB200_ERR_ALLOC;
    {
        LaunchScope ls(ctx, RLB200_TIMER_SKETCH, 2);
        oz_colmax_kernel<<<(unsigned)((N + 7) / 8), 256, 0, ctx->stream>>>(B, ldb, K, (int)N, K, 1, Eb);
        oz_slice_cols_kernel<OZ_BN><<<dim3(nnb, (nkb + 7) / 8), 128, 0, ctx->stream>>>(B, ldb, K, (int)N, nkb, Eb, bt);
        RLB_CUDA_OK(ctx, cudaGetLastError());
    }
    for (int64_t r0 = 0; r0 < m; r0 += RC) {
        const int64_t rows = std::min(RC, m - r0);
        const int nrb = (int)((rows + OZ_BM - 1) / OZ_BM);
        {
            LaunchScope ls(ctx, RLB200_TIMER_SKETCH, 2);
            oz_rowmax_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, ctx->stream>>>(A + r0, lda, rows, (int)K, Ea);
            oz_slice_rows_kernel<OZ_BM><<<dim3(nrb, (nkb + 7) / 8), OZ_BM, 0, ctx->stream>>>(A + r0, lda, rows, (int)K, nkb, Ea, at);
        }
        LaunchScope ls(ctx, RLB200_TIMER_GEMM_NN);
        ozaki_mma_kernel<<<dim3(nnb, nrb, 1), 128, OZ_STAGES * OZ_STAGE_BYTES, ctx->stream>>>(at, 0, bt, 0, nkb, Ea, 0, Eb, 0, rows, (int)N, C + r0, ldc,
                                                                                            0, alpha, beta);
        RLB_CUDA_OK(ctx, cudaGetLastError());
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// C(N1 x N2) = alpha * X(m x N1)^T * Y(m x N2) + beta * C, contraction over the long dimension
// ------------------------------------------------------------------------------------------------
int ozaki_gemm_tn(Ctx* ctx, int64_t m, int64_t N1, int64_t N2, double alpha, const double* X, int64_t ldx, const double* Y, int64_t ldy, double beta,
                  double* C, int64_t ldc) {
    RLB_REQUIRE(ctx, m >= 0 && N1 >= 0 && N2 >= 0 && N1 < (1 << 20) && N2 < (1 << 20));
    if (N1 == 0 || N2 == 0) return 0;
    if (m == 0) return gemm_tn<double>(ctx, 0, N1, N2, 0.0, X, ldx, Y, ldy, beta, C, ldc, 0);
    RLB_CHECK(oz_configure(ctx));
    const int64_t L = std::min<int64_t>(OZ_CHUNK, ((m + OZ_KB - 1) / OZ_KB) * OZ_KB);
    const int64_t nchunks = (m + L - 1) / L;
    const int nkb = (int)(L / OZ_KB);
    const int nb1 = (int)((N1 + OZ_BM - 1) / OZ_BM), nb2 = (int)((N2 + OZ_BN - 1) / OZ_BN);
    // chunks per launch: enough CTAs for ~2 waves
    const int G = (int)std::max<int64_t>(1, std::min<int64_t>(nchunks, (2 * ctx->num_sms + nb1 * nb2 - 1) / (nb1 * nb2)));
    ArenaScope as(ctx);
    const int64_t total = N1 * N2;
    double* part = as.take<double>((size_t)nchunks * total); if (!part) return RLB200_ERR_ALLOC;
    int* Ex = as.take<int>((size_t)nchunks * N1); if (!Ex) return RLB200_ERR_ALLOC;
    int* Ey = as.take<int>((size_t)nchunks * N2); if (!Ey) return RLB200_ERR_ALLOC;
    const int64_t xs = (int64_t)nb1 * nkb * OZ_S * OZ_TILE_A, ys = (int64_t)nb2 * nkb * OZ_S * OZ_TILE_B;   // bytes per chunk
    int8_t* xt = as.take<int8_t>((size_t)G * xs); if (!xt) return RLB200_ERR_ALLOC;
    int8_t* yt = as.take<int8_t>((size_t)G * ys); if (!yt) return RLB200_ERR_ALLOC;
    {
        LaunchScope ls(ctx, RLB200_TIMER_SKETCH, 2);
        oz_colmax_kernel<<<(unsigned)((N1 * nchunks + 7) / 8), 256, 0, ctx->stream>>>(X, ldx, m, (int)N1, L, (int)nchunks, Ex);
        oz_colmax_kernel<<<(unsigned)((N2 * nchunks + 7) / 8), 256, 0, ctx->stream>>>(Y, ldy, m, (int)N2, L, (int)nchunks, Ey);
        RLB_CUDA_OK(ctx, cudaGetLastError());
    }
    for (int64_t c0 = 0; c0 < nchunks; c0 += G) {
        const int g = (int)std::min<int64_t>(G, nchunks - c0);
        {
            LaunchScope ls(ctx, RLB200_TIMER_SKETCH, 2 * g);
            for (int q = 0; q < g; ++q) {
                const int64_t r0 = (c0 + q) * L, klen = std::min(L, m - r0);
                oz_slice_cols_kernel<OZ_BM><<<dim3(nb1, (nkb + 7) / 8), 128, 0, ctx->stream>>>(X + r0, ldx, klen, (int)N1, nkb, Ex + (c0 + q) * N1, xt + q * xs);
                oz_slice_cols_kernel<OZ_BN><<<dim3(nb2, (nkb + 7) / 8), 128, 0, ctx->stream>>>(Y + r0, ldy, klen, (int)N2, nkb, Ey + (c0 + q) * N2, yt + q * ys);
            }
        }
        LaunchScope ls(ctx, RLB200_TIMER_GEMM_TN);
        ozaki_mma_kernel<<<dim3(nb2, nb1, g), 128, OZ_STAGES * OZ_STAGE_BYTES, ctx->stream>>>(xt, xs, yt, ys, nkb, Ex + c0 * N1, N1, Ey + c0 * N2, N2, N1,
                                                                                            (int)N2, part + c0 * total, N1, total, 1.0, 0.0);
        RLB_CUDA_OK(ctx, cudaGetLastError());
    }
    LaunchScope ls(ctx, RLB200_TIMER_GEMM_TN);
    oz_reduce_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(part, (int)nchunks, total, (int)N1,
                                                                                                                       alpha, beta, C, ldc);
    RLB_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

}  // namespace rlb
