// Shared pieces of the int8 digit-slice engine (ozaki.cu: staged digits; ozaki_fused.cu: digits produced inside the tensor-core kernel):
// digit configuration, exponent kernels, slicers, descriptor / mbarrier / bulk-copy / tcgen05 wrappers.
#pragma once
#include "drivers.cuh"

namespace rlb {

constexpr int OZ_KB = 32;          // K bytes (= int8 elements) per stage = one tcgen05.mma K step
constexpr int OZ_BM = 128;         // UMMA M: tile rows of the first operand
constexpr int OZ_BN = 64;          // tile rows of the second operand (per digit)
constexpr int OZ_TILE_A = OZ_BM * OZ_KB;   // bytes of one digit tile of the first operand
constexpr int OZ_TILE_B = OZ_BN * OZ_KB;
constexpr int64_t OZ_CHUNK = 16384;        // rows per int32 accumulation group of the TN product (7 * 2^14 * 2^14 < 2^31)
constexpr int64_t OZ_KMAX = 16384;         // largest K of one NN accumulation group
constexpr int OZ_SMEM_BUDGET = 225 * 1024;

template <int S>
struct OzCfg {
    static_assert(S >= 2 && S <= 7, "digits");
    static constexpr int P = 8 * S - 2;                       // fixed-point bits below the group scale
    static constexpr int STAGE_BYTES = S * (OZ_TILE_A + OZ_TILE_B);
    static constexpr int STAGES = (OZ_SMEM_BUDGET / STAGE_BYTES) > 12 ? 12 : (OZ_SMEM_BUDGET / STAGE_BYTES);
    static constexpr int ACC_COLS = S * OZ_BN;
    static constexpr int TMEM_COLS = ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512);
    // 0x80 in each of the S-1 low bytes: added as a bias it makes those bytes the unsigned digits d + 128, xor-ed it re-centres them
    static constexpr unsigned long long LOWMASK = 0x8080808080808080ull >> (8 * (9 - S));
};

__device__ __forceinline__ uint32_t oz_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------------
// exponents.  Stored value E = max(biased exponent - 1022, P - 1023): |x| < 2^E for the whole group, and 2^(P-E) is a normal
// double (groups whose largest magnitude is below 2^(P-1023) keep fewer digits).
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ int oz_expfield(T x) { return (int)((__double_as_longlong((double)x) >> 52) & 0x7ff); }
__device__ __forceinline__ int oz_exp_from_field(int f, int P) { return max(f - 1022, P - 1023); }

// E_row[i] for rows [0, rows) of A (col-major, lda), K columns.  One thread per row, coalesced across rows.
template <typename T>
__global__ void __launch_bounds__(256) oz_rowexp_kernel(const T* __restrict__ A, int64_t lda, int64_t rows, int K, int P, int* __restrict__ E) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    int f = 0;
    int c = 0;
    for (; c + 4 <= K; c += 4) {
        const T a0 = A[i + (int64_t)c * lda], a1 = A[i + (int64_t)(c + 1) * lda], a2 = A[i + (int64_t)(c + 2) * lda], a3 = A[i + (int64_t)(c + 3) * lda];
        f = max(max(f, oz_expfield(a0)), max(oz_expfield(a1), max(oz_expfield(a2), oz_expfield(a3))));
    }
    for (; c < K; ++c) f = max(f, oz_expfield(A[i + (int64_t)c * lda]));
    E[i] = oz_exp_from_field(f, P);
}
// E[chunk * ncols + c] over rows [chunk*L, (chunk+1)*L) of column c of X (col-major).  One warp per (column, chunk).
// ss (optional): sum of squares of the same entries, same indexing (fixed summation order).
template <typename T>
__global__ void __launch_bounds__(256) oz_colexp_kernel(const T* __restrict__ X, int64_t ldx, int64_t rows, int ncols, int64_t L, int nchunks, int P,
                                                        int* __restrict__ E, double* __restrict__ ss) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= (int64_t)ncols * nchunks) return;
    const int c = (int)(w % ncols), ch = (int)(w / ncols);
    const int64_t r0 = (int64_t)ch * L, r1 = min(rows, r0 + L);
    const T* x = X + (int64_t)c * ldx;
    int f = 0;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int64_t r = r0 + lane;
    for (; r + 96 < r1; r += 128) {
        const double a0 = (double)x[r], a1 = (double)x[r + 32], a2 = (double)x[r + 64], a3 = (double)x[r + 96];
        f = max(max(f, oz_expfield(a0)), max(oz_expfield(a1), max(oz_expfield(a2), oz_expfield(a3))));
        s0 = fma(a0, a0, s0); s1 = fma(a1, a1, s1); s2 = fma(a2, a2, s2); s3 = fma(a3, a3, s3);
    }
    for (; r < r1; r += 32) { const double a0 = (double)x[r]; f = max(f, oz_expfield(a0)); s0 = fma(a0, a0, s0); }
    f = __reduce_max_sync(0xffffffffu, f);
    if (ss) {
        double sv = (s0 + s1) + (s2 + s3);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
        if (lane == 0) ss[(int64_t)ch * ncols + c] = sv;
    }
    if (lane == 0) E[(int64_t)ch * ncols + c] = oz_exp_from_field(f, P);
}
// out[0] = sum of v[0..n) in a fixed order (one CTA)
static __global__ void __launch_bounds__(1024) oz_sum_kernel(const double* __restrict__ v, int64_t n, double* __restrict__ out) {
    __shared__ double sh[1024];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) s += v[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

// Row exponents AND column-chunk exponent fields AND ||A||_F^2 partials of the constant data matrix in ONE sweep (the drivers use A as
// the first operand of both products).  CTA = 256 consecutive rows (never straddling a chunk: L is a multiple of 256), thread = row.
// Column maxima: warp REDUX + one shared-memory atomicMax per warp and column, then one global atomicMax per CTA and column on the
// (pre-zeroed) field array - maxima are order-independent; the sums of squares are combined in a fixed order (lane, warp, CTA).
template <typename T>
__global__ void __launch_bounds__(256) oz_rowcolexp_kernel(const T* __restrict__ A, int64_t lda, int64_t rows, int K, int64_t L, int P_row,
                                                           int* __restrict__ Erow, int* __restrict__ colfield, double* __restrict__ ss_part) {
    extern __shared__ int s_col[];            // [K]
    __shared__ double s_ss[8];
    for (int c = threadIdx.x; c < K; c += 256) s_col[c] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool rv = i < rows;
    const T* a = A + (rv ? i : 0);
    int rowf = 0;
    double ss = 0.0;
    const int lane = threadIdx.x & 31;
    int c = 0;
    for (; c + 4 <= K; c += 4) {
        double x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = rv ? (double)a[(int64_t)(c + u) * lda] : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int f = oz_expfield(x[u]);
            rowf = max(rowf, f);
            ss = fma(x[u], x[u], ss);
            const int wf = __reduce_max_sync(0xffffffffu, f);
            if (lane == 0 && wf > 0) atomicMax(&s_col[c + u], wf);
        }
    }
    for (; c < K; ++c) {
        const double x = rv ? (double)a[(int64_t)c * lda] : 0.0;
        const int f = oz_expfield(x);
        rowf = max(rowf, f);
        ss = fma(x, x, ss);
        const int wf = __reduce_max_sync(0xffffffffu, f);
        if (lane == 0 && wf > 0) atomicMax(&s_col[c], wf);
    }
    if (rv) Erow[i] = oz_exp_from_field(rowf, P_row);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) s_ss[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_ss[w];
        ss_part[blockIdx.x] = t;
    }
    const int64_t chunk = ((int64_t)blockIdx.x * 256) / L;
    for (int cc = threadIdx.x; cc < K; cc += 256)
        if (s_col[cc] > 0) atomicMax(&colfield[chunk * K + cc], s_col[cc]);
}
// fields -> exponents, in place
static __global__ void __launch_bounds__(256) oz_field_to_exp_kernel(int* __restrict__ E, int64_t n, int P) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) E[i] = oz_exp_from_field(E[i], P);
}

// ------------------------------------------------------------------------------------------------
// slicers.  Digit tiles: tile (rb, kb, t) of TR rows x 32 K-bytes at ((rb * nkb + kb) * S + t) * TR * 32, inside it the byte of
// (row r, k) sits at ((r / 8) * 2 + k / 16) * 128 + (r % 8) * 16 + k % 16  (K-major, no swizzle: SBO = 256 B, LBO = 128 B).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double oz_pow2(int e) {      // 2^e for e in [-1022, 1023]
    return __longlong_as_double((long long)(e + 1023) << 52);
}
// F + bias as raw bits: byte b (b < S - 1) is the unsigned digit d + 128 of weight 256^b, byte S - 1 the (two's complement) top digit.
template <int S>
__device__ __forceinline__ unsigned long long oz_fixed(double x, double scale) {
    unsigned long long F;
    if constexpr (OzCfg<S>::P <= 50) {
        // x * scale + 1.5 * 2^52 leaves rn(x * scale) (two's complement) in the low mantissa bits
        F = (unsigned long long)__double_as_longlong(fma(x, scale, 6755399441055744.0)) - 0x4338000000000000ull;
    } else {
        F = (unsigned long long)__double2ll_rn(x * scale);
    }
    return (F + OzCfg<S>::LOWMASK) ^ OzCfg<S>::LOWMASK;
}
// digits of 4 consecutive-k values -> w[t] = the 4 bytes of digit t (t = 0 most significant), k order = byte order
template <int S>
__device__ __forceinline__ void oz_pack4(const unsigned long long f[4], uint32_t* w /* [S] */) {
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { lo[e] = (uint32_t)f[e]; hi[e] = (uint32_t)(f[e] >> 32); }
    uint32_t B[8];
    {
        const uint32_t t0 = __byte_perm(lo[0], lo[1], 0x5140), t1 = __byte_perm(lo[2], lo[3], 0x5140);
        const uint32_t t2 = __byte_perm(lo[0], lo[1], 0x7362), t3 = __byte_perm(lo[2], lo[3], 0x7362);
        B[0] = __byte_perm(t0, t1, 0x5410); B[1] = __byte_perm(t0, t1, 0x7632);
        B[2] = __byte_perm(t2, t3, 0x5410); B[3] = __byte_perm(t2, t3, 0x7632);
    }
    if constexpr (S > 4) {
        const uint32_t t0 = __byte_perm(hi[0], hi[1], 0x5140), t1 = __byte_perm(hi[2], hi[3], 0x5140);
        const uint32_t t2 = __byte_perm(hi[0], hi[1], 0x7362), t3 = __byte_perm(hi[2], hi[3], 0x7362);
        B[4] = __byte_perm(t0, t1, 0x5410); B[5] = __byte_perm(t0, t1, 0x7632);
        B[6] = __byte_perm(t2, t3, 0x5410); B[7] = __byte_perm(t2, t3, 0x7632);
    }
#pragma unroll
    for (int t = 0; t < S; ++t) w[t] = B[S - 1 - t];
}
// The same for 4 values at once, cheaper (S <= 6, i.e. P <= 50): the per-byte bias is folded into the addend of the DFMA (an exact integer
// below 2^40 on top of 1.5 * 2^52), the low 48 bits of the result's bit pattern then ARE F + bias (mod 2^48) - no 64-bit subtraction -, and the
// re-centring xor is applied to the packed words (one LOP3 per digit and 4 values instead of two per value).  Bit-identical to
// oz_fixed + oz_pack4.  scale[e]: power-of-two scale of value e.
template <int S>
__device__ __forceinline__ void oz_convert4(const double* x, const double* scale, uint32_t* w /* [S] */) {
    static_assert(OzCfg<S>::P <= 50, "the fixed-point value must fit below the 1.5 * 2^52 offset");
    const double magic = 6755399441055744.0 + (double)OzCfg<S>::LOWMASK;
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const double t = fma(x[e], scale[e], magic);
        lo[e] = (uint32_t)__double2loint(t);
        hi[e] = (uint32_t)__double2hiint(t);
    }
    uint32_t B[8];
    {
        const uint32_t t0 = __byte_perm(lo[0], lo[1], 0x5140), t1 = __byte_perm(lo[2], lo[3], 0x5140);
        const uint32_t t2 = __byte_perm(lo[0], lo[1], 0x7362), t3 = __byte_perm(lo[2], lo[3], 0x7362);
        B[0] = __byte_perm(t0, t1, 0x5410); B[1] = __byte_perm(t0, t1, 0x7632);
        B[2] = __byte_perm(t2, t3, 0x5410); B[3] = __byte_perm(t2, t3, 0x7632);
    }
    if constexpr (S > 4) {
        const uint32_t t0 = __byte_perm(hi[0], hi[1], 0x5140), t1 = __byte_perm(hi[2], hi[3], 0x5140);
        B[4] = __byte_perm(t0, t1, 0x5410); B[5] = __byte_perm(t0, t1, 0x7632);
    }
#pragma unroll
    for (int t = 0; t < S; ++t) w[t] = (t == 0) ? B[S - 1] : (B[S - 1 - t] ^ 0x80808080u);
}

// 16 consecutive-k values of one tile row -> one 16-byte chunk per digit
template <int S>
__device__ __forceinline__ void oz_emit16(const double* xv, double scale, int8_t* tile0, int64_t tile_bytes, int off) {
    uint32_t pk[4][S];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if constexpr (OzCfg<S>::P <= 50) {
            const double sv[4] = {scale, scale, scale, scale};
            oz_convert4<S>(xv + 4 * q, sv, pk[q]);
        } else {
            unsigned long long f[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) f[e] = oz_fixed<S>(xv[4 * q + e], scale);
            oz_pack4<S>(f, pk[q]);
        }
    }
#pragma unroll
    for (int t = 0; t < S; ++t)
        *reinterpret_cast<uint4*>(tile0 + (int64_t)t * tile_bytes + off) = make_uint4(pk[0][t], pk[1][t], pk[2][t], pk[3][t]);
}

// First operand of the NN product: tile rows = rows of A, K = columns of A, scale per row.  CTA = (row block, group of 8 K-blocks).
template <int S, typename T>
__global__ void __launch_bounds__(OZ_BM) oz_slice_rows_kernel(const T* __restrict__ A, int64_t lda, int64_t rows, int K, int nkb,
                                                              const int* __restrict__ E, int8_t* __restrict__ out) {
    constexpr int TR = OZ_BM;
    const int r = threadIdx.x;
    const int64_t rb = blockIdx.x;
    const int64_t row = rb * TR + r;
    const bool rv = row < rows;
    const double scale = rv ? oz_pow2(OzCfg<S>::P - E[row]) : 0.0;
    const T* a = A + (rv ? row : 0);
    for (int kb = blockIdx.y * 8; kb < min(nkb, (int)blockIdx.y * 8 + 8); ++kb) {
        int8_t* tile0 = out + ((rb * nkb + kb) * S) * (int64_t)(TR * OZ_KB);
        T raw[32];
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) {
            const int col = kb * OZ_KB + kk;
            raw[kk] = (rv && col < K) ? a[(int64_t)col * lda] : T(0);
        }
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
            double xv[16];
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) xv[kk] = (double)raw[kc * 16 + kk];
            oz_emit16<S>(xv, scale, tile0, TR * OZ_KB, ((r >> 3) * 2 + kc) * 128 + (r & 7) * 16);
        }
    }
}

// Operands whose K runs along the contiguous direction: tile rows = columns of X, K = rows [0, klen) of X, scale per column
// (E[c], already offset to the chunk).  CTA = (column block, group of 8 K-blocks); thread = (column, 16-element K chunk).
template <int S, int TR, typename T>
__global__ void __launch_bounds__(128) oz_slice_cols_kernel(const T* __restrict__ X, int64_t ldx, int64_t klen, int ncols, int nkb,
                                                            const int* __restrict__ E, int8_t* __restrict__ out) {
    const int64_t cb = blockIdx.x;
    for (int kb = blockIdx.y * 8; kb < min(nkb, (int)blockIdx.y * 8 + 8); ++kb) {
        int8_t* tile0 = out + ((cb * nkb + kb) * S) * (int64_t)(TR * OZ_KB);
        for (int item = threadIdx.x; item < TR * 2; item += 128) {
            const int cl = item >> 1, kc = item & 1;
            const int64_t c = cb * TR + cl;
            const bool cv = c < ncols;
            const double scale = cv ? oz_pow2(OzCfg<S>::P - E[c]) : 0.0;
            const int64_t kbase = (int64_t)kb * OZ_KB + kc * 16;
            const T* x = X + (cv ? c : 0) * ldx + kbase;
            double xv[16];
            if (cv && kbase + 16 <= klen && ((reinterpret_cast<uintptr_t>(x) & 15) == 0)) {
                if constexpr (sizeof(T) == 8) {
#pragma unroll
                    for (int kk = 0; kk < 16; kk += 2) { const double2 t2 = *reinterpret_cast<const double2*>(x + kk); xv[kk] = t2.x; xv[kk + 1] = t2.y; }
                } else {
#pragma unroll
                    for (int kk = 0; kk < 16; kk += 4) {
                        const float4 t4 = *reinterpret_cast<const float4*>(x + kk);
                        xv[kk] = t4.x; xv[kk + 1] = t4.y; xv[kk + 2] = t4.z; xv[kk + 3] = t4.w;
                    }
                }
            } else {
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) xv[kk] = (cv && kbase + kk < klen) ? (double)x[kk] : 0.0;
            }
            oz_emit16<S>(xv, scale, tile0, TR * OZ_KB, ((cl >> 3) * 2 + kc) * 128 + (cl & 7) * 16);
        }
    }
}

// TN operands, MN-major tiles: tile rows (MN) = columns of X, K = rows [0, klen) of X, scale per column (E[c], already offset to
// the chunk).  Inside a tile the byte of (column c, row k) sits at ((c / 16) * 4 + k / 8) * 128 + (k % 8) * 16 + c % 16, i.e. the
// same 8 x 16-byte core matrix as above read the other way round (LBO = 128 B between 8-row K groups, SBO = 512 B between 16-column
// groups).  This lets a thread own one ROW: a warp reads 32 consecutive rows of a column (256 contiguous bytes per load) and writes 512
// contiguous bytes per digit - the access pattern of the NN slicer, which the K-major column slicer above cannot have.
// CTA = 4 warps = 4 consecutive K-blocks of one column block.
template <int S, int TR, typename T>
__global__ void __launch_bounds__(128) oz_slice_tn_kernel(const T* __restrict__ X, int64_t ldx, int64_t klen, int ncols, int nkb,
                                                          const int* __restrict__ E, int8_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int kb = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (kb >= nkb) return;
    const int64_t cb = blockIdx.y;
    const int64_t row = (int64_t)kb * OZ_KB + lane;
    const bool rv = row < klen;
    const T* x = X + (rv ? row : 0);
    int8_t* tile0 = out + ((cb * nkb + kb) * S) * (int64_t)(TR * OZ_KB);
#pragma unroll 1
    for (int cg = 0; cg < TR / 16; ++cg) {
        const int64_t c0 = cb * TR + cg * 16;
        double xv[16];
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) xv[kk] = (rv && c0 + kk < ncols) ? (double)x[(c0 + kk) * ldx] : 0.0;
        uint32_t pk[4][S];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double sv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int64_t c = c0 + 4 * q + e;
                sv[e] = oz_pow2(OzCfg<S>::P - E[c < ncols ? c : 0]);
            }
            if constexpr (OzCfg<S>::P <= 50) {
                oz_convert4<S>(xv + 4 * q, sv, pk[q]);
            } else {
                unsigned long long f[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) f[e] = oz_fixed<S>(xv[4 * q + e], sv[e]);
                oz_pack4<S>(f, pk[q]);
            }
        }
        const int off = (cg * 4 + (lane >> 3)) * 128 + (lane & 7) * 16;
#pragma unroll
        for (int t = 0; t < S; ++t)
            *reinterpret_cast<uint4*>(tile0 + (int64_t)t * (TR * OZ_KB) + off) = make_uint4(pk[0][t], pk[1][t], pk[2][t], pk[3][t]);
    }
}

// ------------------------------------------------------------------------------------------------
// the tensor-core kernel
// ------------------------------------------------------------------------------------------------
// SWIZZLE_NONE shared-memory matrix descriptor (version 1); core matrix = 8 x 16 bytes = 128 contiguous bytes.
//   K-major  (MN = false): 8 tile rows x 16 K-bytes;  LBO = 128 B between the two K halves, SBO = 256 B between 8-row groups
//   MN-major (MN = true):  16 MN-bytes x 8 K-rows;    LBO = 128 B between 8-row K groups,   SBO = 512 B between 16-wide MN groups
template <bool MN>
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr) {
    constexpr uint64_t SBO = MN ? 512 : 256;
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((SBO >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void oz_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
// the same with a suspend-time hint (ns): the thread is woken when the phase completes, the hint only bounds how long the hardware may keep
// it suspended before the instruction returns false - fewer re-polls of the barrier word through the shared-memory pipe
__device__ __forceinline__ void oz_mbar_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(ns) : "memory");
    }
}
// non-blocking: has the phase completed?
__device__ __forceinline__ bool oz_mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ bool oz_mbar_test_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
// acquire at cluster scope: the data guarded by the barrier may have been written by a peer CTA (distributed shared memory)
__device__ __forceinline__ void oz_mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void oz_mbar_wait_cluster_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(ns) : "memory");
    }
}
// long waits (the epilogue warps wait for a whole tile's main loop): back off so that the spinning warps do not take issue slots from
// the slicer CTAs that share the SM
__device__ __forceinline__ void oz_mbar_wait_sleep(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(256);
    }
}
__device__ __forceinline__ void oz_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void oz_bulk_load_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void oz_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void oz_mma_i8(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// v = 2^16 * sum_d acc_d 256^-d from the S int32 accumulators of one output (acc[d] = the raw TMEM word of anti-diagonal d).
// Each accumulator becomes a double exactly with one LOP3 + one DADD (2^52 + 2^31 trick; no I2F.F64), then Horner in fp64: the first two
// steps are exact (<= 47 bits), the rest round at 2^-53 of the running sum - far below the 2^-(8S-3) of the scheme.  12 fp64 + 6 integer
// instructions per output for S = 6; the previous exact int64 combination cost ~32 integer + 4 fp64 and made the epilogue ALU-bound
// (10.6 k cycles per 128 x 64 tile; the tensor pipe idles meanwhile).  Fixed evaluation order: results are reproducible run to run.
template <int S>
__device__ __forceinline__ double oz_combine(const uint32_t* acc /* [S], stride 8 words between diagonals */) {
    auto cvt = [](uint32_t a) { return __hiloint2double(0x43300000, (int)(a ^ 0x80000000u)) - 4503601774854144.0; };
    double h = cvt(acc[(S - 1) * 8]);
#pragma unroll
    for (int d = S - 2; d >= 0; --d) h = fma(h, 0.00390625, cvt(acc[d * 8]));
    return h * 65536.0;
}

// C = alpha * sum_g part[g] + beta * C  (fixed order)
template <typename T>
__global__ void __launch_bounds__(256) oz_reduce_kernel(const double* __restrict__ part, int groups, int64_t total, int n1, double alpha, double beta,
                                                        T* __restrict__ C, int64_t ldc) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int g = 0; g < groups; ++g) s += part[(int64_t)g * total + e];
        T* c = C + (e % n1) + (e / n1) * ldc;
        double v = alpha * s;
        if (beta != 0.0) v += beta * (double)(*c);
        *c = (T)v;
    }
}

}  // namespace rlb
