// Tall-skinny GEMM family on the fp64 tensor pipe (DMMA.8x8x4), cp.async multi-stage smem pipeline.
//
//   gemm_nn : C(m x N) = alpha * A(m x K) * B(K x N) + beta * C      long dimension = m (rows)
//             replaces blas::gemm(NoTrans,NoTrans) at rl_rs.hh:153, rl_rf.hh:123, rl_rsvd.hh:148 and the
//             trsm at rl_orth.hh:95 (as a product with the explicit inverse of R).
//             May run IN PLACE (C == A) when one CTA owns all N columns of its rows (TN >= N).
//   gemm_tn : C(N1 x N2) = alpha * A(m x N1)^T * B(m x N2) + beta * C  long dimension = m (contraction)
//             split-K over m with a deterministic two-stage reduction; `upper_only` skips tiles strictly
//             below the diagonal (syrk, rl_orth.hh:78).  Replaces blas::gemm(Trans,NoTrans) at
//             rl_rs.hh:142,165 and rl_qb.hh:218.
//
// Storage type T in {double, float}; arithmetic is always fp64 (float inputs are widened when the
// fragments are read from shared memory), so the float path is at least as accurate as the reference's.
// Roofline: both kernels are bound by the fp64 tensor/FMA pipe (2*m*N*K flop; intensity N/4..K/4 flop/B).
#include "common.cuh"

namespace rlb {

template <typename T> struct Pad;
template <> struct Pad<double> { static constexpr int A = 4, B = 4; };   // strides = 4 (mod 16) 8-byte words
template <> struct Pad<float>  { static constexpr int A = 8, B = 4; };   // strides = 8 / 4 (mod 32) 4-byte words

// ------------------------------------------------------------------------------------------------
// NN
// ------------------------------------------------------------------------------------------------
template <typename T, int TM, int TN, int WGM, int WGN, int KS, int STAGES>
struct NNCfg {
    static constexpr int THREADS = WGM * WGN * 32;
    static constexpr int WM = TM / WGM, WN = TN / WGN;      // warp tile
    static constexpr int MI = WM / 8, NI = WN / 8;          // 8x8 mma tiles per warp
    static constexpr int SA = TM + Pad<T>::A;               // sA[k][row]
    static constexpr int SB = KS + Pad<T>::B;               // sB[n][k]
    static constexpr int A_ELEMS = KS * SA, B_ELEMS = TN * SB;
    static constexpr size_t SMEM = (size_t)STAGES * (A_ELEMS + B_ELEMS) * sizeof(T);
};

// TB : B is given transposed (N x K).   TRI: B is upper triangular (B[k][n] = 0 for k > n): MMA column tiles are dealt to the
// warps round-robin and every 8-column tile skips the k-range below its diagonal (~half the work, evenly spread).
template <typename T, int TM, int TN, int WGM, int WGN, int KS, int STAGES, int MINB, bool TB, bool TRI>
__global__ void __launch_bounds__(WGM* WGN * 32, MINB)
gemm_nn_kernel(int64_t m, int N, int K, double alpha, const T* __restrict__ A, int64_t lda, const T* __restrict__ B, int64_t ldb,
               double beta, T* C, int64_t ldc, int a_al16, int b_al16) {
    using Cfg = NNCfg<T, TM, TN, WGM, WGN, KS, STAGES>;
    constexpr int E = 16 / sizeof(T);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sA = reinterpret_cast<T*>(smem_raw);
    T* sB = sA + (size_t)STAGES * Cfg::A_ELEMS;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp % WGM, wn = warp / WGM;
    // 1-D grid, N-tile index fastest: the CTAs that share a row block of A are launched back to back (L2 reuse)
    const int ntiles = (N + TN - 1) / TN;
    const int64_t m0 = (int64_t)(blockIdx.x / ntiles) * TM;
    const int n0 = (int)(blockIdx.x % ntiles) * TN;
    const int nk = (K + KS - 1) / KS;

    auto load_stage = [&](int stage, int kt) {
        const int k0 = kt * KS;
        T* a = sA + (size_t)stage * Cfg::A_ELEMS;
        T* b = sB + (size_t)stage * Cfg::B_ELEMS;
        // A tile: KS columns x TM rows, chunks along rows
        constexpr int ACH = TM / E;
        for (int c = tid; c < KS * ACH; c += Cfg::THREADS) {
            const int k = c / ACH, r = (c % ACH) * E;
            const int64_t gr = m0 + r;
            int valid = (k0 + k < K) ? (int)min((int64_t)E, m - gr) : 0;
            load_chunk<T>(a + k * Cfg::SA + r, A + gr + (int64_t)(k0 + k) * lda, valid, a_al16);
        }
        if (!TB) {
            // B tile: TN columns x KS rows, chunks along k
            constexpr int BCH = KS / E;
            for (int c = tid; c < TN * BCH; c += Cfg::THREADS) {
                const int n = c / BCH, k = (c % BCH) * E;
                int valid = (n0 + n < N) ? min(E, K - (k0 + k)) : 0;
                load_chunk<T>(b + n * Cfg::SB + k, B + (k0 + k) + (int64_t)(n0 + n) * ldb, valid, b_al16);
            }
        } else {
            // B given transposed (N x K, ld = ldb): element-wise transpose into sB[n][k]
            for (int c = tid; c < TN * KS; c += Cfg::THREADS) {
                const int k = c / TN, n = c % TN;
                const bool v = (n0 + n < N) && (k0 + k < K);
                const T* src = v ? B + (n0 + n) + (int64_t)(k0 + k) * ldb : B;
                if (sizeof(T) == 8) cp_async_8(b + n * Cfg::SB + k, src, v);
                else                cp_async_4(b + n * Cfg::SB + k, src, v);
            }
        }
    };

    double acc[Cfg::MI][Cfg::NI][2];
#pragma unroll
    for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
        for (int j = 0; j < Cfg::NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }

    const int arow = wm * Cfg::WM + (lane >> 2), ak = lane & 3;
    const int bk = lane & 3;
    // first column (within the CTA tile) of this warp's j-th 8-column MMA tile
    auto tile_col = [&](int j) { return TRI ? (j * WGN + wn) * 8 : wn * Cfg::WN + j * 8; };

    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {   // prefetch tile kt+STAGES-1 into the slot freed by iteration kt-1
            const int nxt = kt + STAGES - 1;
            if (nxt < nk) load_stage(nxt % STAGES, nxt);
            cp_async_commit();
        }
        const T* a = sA + (size_t)(kt % STAGES) * Cfg::A_ELEMS;
        const T* b = sB + (size_t)(kt % STAGES) * Cfg::B_ELEMS;
#pragma unroll
        for (int kk = 0; kk < KS; kk += 4) {
            double af[Cfg::MI], bf[Cfg::NI];
#pragma unroll
            for (int i = 0; i < Cfg::MI; ++i) af[i] = (double)a[(kk + ak) * Cfg::SA + arow + i * 8];
#pragma unroll
            for (int j = 0; j < Cfg::NI; ++j) bf[j] = (double)b[(tile_col(j) + (lane >> 2)) * Cfg::SB + kk + bk];
#pragma unroll
            for (int j = 0; j < Cfg::NI; ++j) {
                // rows k >= (last column of the tile) + 1 of an upper-triangular B are zero (warp-uniform test)
                if (TRI && kt * KS + kk >= n0 + tile_col(j) + 8) continue;
#pragma unroll
                for (int i = 0; i < Cfg::MI; ++i) dmma_884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();   // every warp is past its last smem read (and, for in-place use, every read of A is done)

    // epilogue
    const int crow = wm * Cfg::WM + (lane >> 2);
#pragma unroll
    for (int i = 0; i < Cfg::MI; ++i) {
        const int64_t gr = m0 + crow + i * 8;
        if (gr >= m) continue;
#pragma unroll
        for (int j = 0; j < Cfg::NI; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int gc = n0 + tile_col(j) + 2 * (lane & 3) + h;
                if (gc < N) {
                    T* p = C + gr + (int64_t)gc * ldc;
                    double v = alpha * acc[i][j][h];
                    if (beta != 0.0) v += beta * (double)(*p);
                    *p = (T)v;
                }
            }
        }
    }
}

template <typename T>
static bool al16(const T* p, int64_t ld) {
    constexpr int E = 16 / sizeof(T);
    return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (ld % E == 0);
}

template <typename T, int TM, int TN, int WGM, int WGN, int KS, int STAGES, int MINB, bool TB = false, bool TRI = false>
int launch_nn(Ctx* ctx, int64_t m, int N, int K, double alpha, const T* A, int64_t lda, const T* B, int64_t ldb, double beta,
                     T* C, int64_t ldc) {
    using Cfg = NNCfg<T, TM, TN, WGM, WGN, KS, STAGES>;
    auto kern = gemm_nn_kernel<T, TM, TN, WGM, WGN, KS, STAGES, MINB, TB, TRI>;
    RLB_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    const int64_t nblk = ((m + TM - 1) / TM) * ((N + TN - 1) / TN);
    RLB_REQUIRE(ctx, nblk < (1ll << 31));
    dim3 grid((unsigned)nblk);
    LaunchScope ls(ctx, RLB200_TIMER_GEMM_NN);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, ctx->stream>>>(m, N, K, alpha, A, lda, B, ldb, beta, C, ldc, al16(A, lda), TB ? 0 : al16(B, ldb));
    RLB_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

// C = alpha*A*B + beta*C, tall A.  In-place (C == A, requires K == N <= 256... handled by caller via gemm_nn_inplace).
template <typename T>
int gemm_nn(Ctx* ctx, int64_t m, int64_t N, int64_t K, double alpha, const T* A, int64_t lda, const T* B, int64_t ldb, double beta,
            T* C, int64_t ldc) {
    RLB_REQUIRE(ctx, m >= 0 && N >= 0 && K >= 0 && N < (1ll << 30) && K < (1ll << 30));
    RLB_REQUIRE(ctx, (m + 63) / 64 < (1ll << 31));
    if (m == 0 || N == 0) return 0;
    // tile shapes from the sweep in tools/gemm_tune.cu (profiles/gemm_tune_r1.json): two CTAs per SM hide each other's
    // barriers, prologue and epilogue; 128x64x32 reaches 89% of the measured DMMA peak on the C2 shape
    if (N <= 32)  return launch_nn<T, 128, 32, 4, 1, 16, 4, 2>(ctx, m, (int)N, (int)K, alpha, A, lda, B, ldb, beta, C, ldc);
    return launch_nn<T, 128, 64, 4, 2, 32, 2, 2>(ctx, m, (int)N, (int)K, alpha, A, lda, B, ldb, beta, C, ldc);
}

// C = alpha*A*B^T + beta*C with B stored N x K (the deflation update of rl_qb.hh:260)
template <typename T>
int gemm_nt(Ctx* ctx, int64_t m, int64_t N, int64_t K, double alpha, const T* A, int64_t lda, const T* B, int64_t ldb, double beta,
            T* C, int64_t ldc) {
    RLB_REQUIRE(ctx, m >= 0 && N >= 0 && K >= 0 && N < (1ll << 30) && K < (1ll << 30));
    if (m == 0 || N == 0) return 0;
    return launch_nn<T, 128, 64, 4, 2, 32, 2, 2, true>(ctx, m, (int)N, (int)K, alpha, A, lda, B, ldb, beta, C, ldc);
}

// X <- alpha * X * B  in place (X: m x K, B: K x N with N <= K so the result fits in X's columns 0..N-1).
// One CTA owns every column of its 64 rows, so all of X's tile is consumed before anything is stored.
template <typename T>
int gemm_nn_inplace(Ctx* ctx, int64_t m, int64_t N, int64_t K, double alpha, T* X, int64_t ldx, const T* B, int64_t ldb, bool b_upper_tri) {
    RLB_REQUIRE(ctx, N <= 256 && N <= K);
    if (m == 0 || N == 0) return 0;
    if (b_upper_tri) {
        if (N <= 64)  return launch_nn<T, 128, 64, 4, 2, 32, 2, 2, false, true>(ctx, m, (int)N, (int)K, alpha, X, ldx, B, ldb, 0.0, X, ldx);
        if (N <= 128) return launch_nn<T, 128, 128, 2, 4, 32, 3, 1, false, true>(ctx, m, (int)N, (int)K, alpha, X, ldx, B, ldb, 0.0, X, ldx);
        return launch_nn<T, 64, 256, 2, 4, 32, 2, 1, false, true>(ctx, m, (int)N, (int)K, alpha, X, ldx, B, ldb, 0.0, X, ldx);
    }
    if (N <= 64)  return launch_nn<T, 128, 64, 4, 2, 32, 2, 2>(ctx, m, (int)N, (int)K, alpha, X, ldx, B, ldb, 0.0, X, ldx);
    if (N <= 128) return launch_nn<T, 128, 128, 2, 4, 32, 3, 1>(ctx, m, (int)N, (int)K, alpha, X, ldx, B, ldb, 0.0, X, ldx);
    return launch_nn<T, 64, 256, 2, 4, 32, 2, 1>(ctx, m, (int)N, (int)K, alpha, X, ldx, B, ldb, 0.0, X, ldx);
}

// ------------------------------------------------------------------------------------------------
// TN split-K
// ------------------------------------------------------------------------------------------------
template <typename T, int T1, int T2, int WG1, int WG2, int KS, int STAGES>
struct TNCfg {
    static constexpr int THREADS = WG1 * WG2 * 32;
    static constexpr int W1 = T1 / WG1, W2 = T2 / WG2;
    static constexpr int MI = W1 / 8, NI = W2 / 8;
    static constexpr int SK = KS + Pad<T>::B;               // sA[i][r], sB[j][r]
    static constexpr int A_ELEMS = T1 * SK, B_ELEMS = T2 * SK;
    static constexpr size_t SMEM = (size_t)STAGES * (A_ELEMS + B_ELEMS) * sizeof(T);
};

// partial[split][j][i] (column-major N1 x N2 per split, ld = N1) = A[rows of split]^T B[rows of split]
template <typename T, int T1, int T2, int WG1, int WG2, int KS, int STAGES, int MINB>
__global__ void __launch_bounds__(WG1* WG2 * 32, MINB)
gemm_tn_kernel(int64_t m, int N1, int N2, const T* __restrict__ A, int64_t lda, const T* __restrict__ B, int64_t ldb,
               double* __restrict__ partial, int64_t rows_per_split, int tiles1, int upper_only, int a_al16, int b_al16,
               double* __restrict__ sq_part) {
    using Cfg = TNCfg<T, T1, T2, WG1, WG2, KS, STAGES>;
    constexpr int E = 16 / sizeof(T);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sA = reinterpret_cast<T*>(smem_raw);
    T* sB = sA + (size_t)STAGES * Cfg::A_ELEMS;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int w1 = warp % WG1, w2 = warp / WG1;
    const int t1 = blockIdx.x % tiles1, t2 = blockIdx.x / tiles1;
    const int split = blockIdx.y;
    const int i0 = t1 * T1, j0 = t2 * T2;
    if (upper_only && i0 >= j0 + T2) return;   // tile strictly below the diagonal
    const int64_t r_begin = (int64_t)split * rows_per_split;
    const int64_t r_end = min(m, r_begin + rows_per_split);
    const int nk = r_end > r_begin ? (int)((r_end - r_begin + KS - 1) / KS) : 0;

    auto load_stage = [&](int stage, int kt) {
        const int64_t r0 = r_begin + (int64_t)kt * KS;
        T* a = sA + (size_t)stage * Cfg::A_ELEMS;
        T* b = sB + (size_t)stage * Cfg::B_ELEMS;
        constexpr int CH = KS / E;
        for (int c = tid; c < T1 * CH; c += Cfg::THREADS) {
            const int i = c / CH, r = (c % CH) * E;
            int valid = (i0 + i < N1) ? (int)min((int64_t)E, r_end - (r0 + r)) : 0;
            load_chunk<T>(a + i * Cfg::SK + r, A + (r0 + r) + (int64_t)(i0 + i) * lda, valid, a_al16);
        }
        for (int c = tid; c < T2 * CH; c += Cfg::THREADS) {
            const int j = c / CH, r = (c % CH) * E;
            int valid = (j0 + j < N2) ? (int)min((int64_t)E, r_end - (r0 + r)) : 0;
            load_chunk<T>(b + j * Cfg::SK + r, B + (r0 + r) + (int64_t)(j0 + j) * ldb, valid, b_al16);
        }
    };

    double acc[Cfg::MI][Cfg::NI][2];
#pragma unroll
    for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
        for (int j = 0; j < Cfg::NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    const int arow = w1 * Cfg::W1 + (lane >> 2), bcol = w2 * Cfg::W2 + (lane >> 2), kq = lane & 3;
    const bool do_sq = (sq_part != nullptr) && (t2 == 0) && (w2 == 0);
    const bool cta_sq = (sq_part != nullptr) && (t2 == 0);
    double sq = 0.0;
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nxt = kt + STAGES - 1;
            if (nxt < nk) load_stage(nxt % STAGES, nxt);
            cp_async_commit();
        }
        const T* a = sA + (size_t)(kt % STAGES) * Cfg::A_ELEMS;
        const T* b = sB + (size_t)(kt % STAGES) * Cfg::B_ELEMS;
#pragma unroll
        for (int kk = 0; kk < KS; kk += 4) {
            double af[Cfg::MI], bf[Cfg::NI];
#pragma unroll
            for (int i = 0; i < Cfg::MI; ++i) af[i] = (double)a[(arow + i * 8) * Cfg::SK + kk + kq];
#pragma unroll
            for (int j = 0; j < Cfg::NI; ++j) bf[j] = (double)b[(bcol + j * 8) * Cfg::SK + kk + kq];
            // fused lange(Fro, A) (rl_qb.hh:168): the warps of the first column group hold every element of the A tile exactly once
            // in their MMA fragments (zero outside the matrix), so the sum of squares costs MI DFMAs per 16 DMMAs and no extra loads
            if (do_sq) {
#pragma unroll
                for (int i = 0; i < Cfg::MI; ++i) sq = fma(af[i], af[i], sq);
            }
#pragma unroll
            for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
                for (int j = 0; j < Cfg::NI; ++j) dmma_884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();
    if (cta_sq) {
        __shared__ double sq_sh[32];
        sq = warp_sum(sq);
        if (lane == 0) sq_sh[warp] = sq;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < Cfg::THREADS / 32; ++w) t += sq_sh[w];
            sq_part[(int64_t)split * tiles1 + t1] = t;
        }
    }

    double* P = partial + (int64_t)split * N1 * N2;
    const int crow = w1 * Cfg::W1 + (lane >> 2), ccol = w2 * Cfg::W2 + 2 * (lane & 3);
#pragma unroll
    for (int i = 0; i < Cfg::MI; ++i) {
        const int gi = i0 + crow + i * 8;
        if (gi >= N1) continue;
#pragma unroll
        for (int j = 0; j < Cfg::NI; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int gj = j0 + ccol + j * 8 + h;
                if (gj < N2) P[gi + (int64_t)gj * N1] = acc[i][j][h];
            }
    }
}

// out = sum(part[0..n)) in a fixed order (single warp)
__global__ void sum_fixed_order_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) s += part[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) *out = s;
}

// C = alpha * sum_s partial[s] + beta * C ; fixed summation order => run-to-run deterministic
template <typename T>
__global__ void splitk_reduce_kernel(const double* __restrict__ partial, int splits, int N1, int N2, double alpha, double beta, T* C,
                                     int64_t ldc, int upper_only, int T1, int T2) {
    const int64_t total = (int64_t)N1 * N2;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e % N1), j = (int)(e / N1);
        if (upper_only && (i / T1) * T1 >= (j / T2) * T2 + T2) continue;   // tile never computed
        double s = 0.0;
        for (int p = 0; p < splits; ++p) s += partial[(int64_t)p * total + e];
        T* c = C + i + (int64_t)j * ldc;
        double v = alpha * s;
        if (beta != 0.0) v += beta * (double)(*c);
        *c = (T)v;
    }
}

template <typename T, int T1, int T2, int WG1, int WG2, int KS, int STAGES, int MINB = 1>
int launch_tn(Ctx* ctx, int64_t m, int N1, int N2, double alpha, const T* A, int64_t lda, const T* B, int64_t ldb, double beta,
                     T* C, int64_t ldc, int upper_only, double* a_sumsq_out = nullptr) {
    using Cfg = TNCfg<T, T1, T2, WG1, WG2, KS, STAGES>;
    const int tiles1 = (N1 + T1 - 1) / T1, tiles2 = (N2 + T2 - 1) / T2;
    const int tiles = tiles1 * tiles2;
    // tiles that actually run (syrk skips the ones strictly below the diagonal)
    int active = 0;
    for (int t2 = 0; t2 < tiles2; ++t2)
        for (int t1 = 0; t1 < tiles1; ++t1) active += !(upper_only && t1 * T1 >= t2 * T2 + T2);
    // split-K factor: every CTA does the same amount of work, so pick the smallest number of splits that makes the number of
    // working CTAs a whole number of waves (a multiple of the SM count; 1 CTA per SM), then cap by the matrix height.
    auto gcd = [](int64_t a, int64_t b) { while (b) { int64_t t = a % b; a = b; b = t; } return a; };
    const int64_t slots = (int64_t)ctx->num_sms * MINB;   // CTAs resident at once
    int64_t want = slots / gcd(active, slots);
    while (want * active < 2 * slots) want *= 2;
    if (active >= 4 * slots) want = 1;   // enough tiles to fill the machine several times over: no split, no partial-sum traffic
    const int64_t max_splits = std::max<int64_t>(1, m / (8 * KS));
    want = std::min(want, max_splits);
    int64_t rows_per_split = (m + want - 1) / want;
    rows_per_split = std::max<int64_t>(((rows_per_split + KS - 1) / KS) * KS, KS);
    const int splits = (int)std::max<int64_t>(1, (m + rows_per_split - 1) / rows_per_split);
    const size_t pbytes = (size_t)splits * N1 * N2 * sizeof(double);
    RLB_CHECK(ws_reserve(ctx, pbytes + (size_t)splits * tiles1 * sizeof(double)));
    double* partial = static_cast<double*>(ctx->ws);
    double* sq_part = a_sumsq_out ? partial + (size_t)splits * N1 * N2 : nullptr;
    auto kern = gemm_tn_kernel<T, T1, T2, WG1, WG2, KS, STAGES, MINB>;
    RLB_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    LaunchScope ls(ctx, RLB200_TIMER_GEMM_TN, a_sumsq_out ? 3 : 2);
    dim3 grid((unsigned)tiles, (unsigned)splits);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, ctx->stream>>>(m, N1, N2, A, lda, B, ldb, partial, rows_per_split, tiles1, upper_only,
                                                         al16(A, lda), al16(B, ldb), sq_part);
    RLB_CUDA_OK(ctx, cudaGetLastError());
    if (a_sumsq_out) sum_fixed_order_kernel<<<1, 32, 0, ctx->stream>>>(sq_part, splits * tiles1, a_sumsq_out);
    const int64_t total = (int64_t)N1 * N2;
    int rb = (int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->num_sms * 8);
    splitk_reduce_kernel<T><<<rb, 256, 0, ctx->stream>>>(partial, splits, N1, N2, alpha, beta, C, ldc, upper_only, T1, T2);
    RLB_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

// NOTE: uses ctx->ws for the split-K partials; callers must not hold live data at the start of ctx->ws.
template <typename T>
int gemm_tn(Ctx* ctx, int64_t m, int64_t N1, int64_t N2, double alpha, const T* A, int64_t lda, const T* B, int64_t ldb, double beta,
            T* C, int64_t ldc, int upper_only, double* a_sumsq_out) {
    RLB_REQUIRE(ctx, m >= 0 && N1 >= 0 && N2 >= 0 && N1 < (1 << 20) && N2 < (1 << 20));
    if (N1 == 0 || N2 == 0) return 0;
    // tile shapes from tools/gemm_tune.cu: 64x128x32 with 2 CTAs/SM reaches 90% of the DMMA peak on A^T*Y (C2 shape);
    // the Gram (syrk) case prefers 64x64 tiles (finer triangle, 4 CTAs/SM)
    if (upper_only || (N1 <= 64 && N2 <= 64))
        return launch_tn<T, 64, 64, 2, 2, 16, 3, 4>(ctx, m, (int)N1, (int)N2, alpha, A, lda, B, ldb, beta, C, ldc, upper_only, a_sumsq_out);
    return launch_tn<T, 64, 128, 2, 4, 32, 2, 2>(ctx, m, (int)N1, (int)N2, alpha, A, lda, B, ldb, beta, C, ldc, upper_only, a_sumsq_out);
}

#define INST(T)                                                                                                                  \
    template int gemm_nn<T>(Ctx*, int64_t, int64_t, int64_t, double, const T*, int64_t, const T*, int64_t, double, T*, int64_t);  \
    template int gemm_nt<T>(Ctx*, int64_t, int64_t, int64_t, double, const T*, int64_t, const T*, int64_t, double, T*, int64_t);  \
    template int gemm_nn_inplace<T>(Ctx*, int64_t, int64_t, int64_t, double, T*, int64_t, const T*, int64_t, bool);                     \
    template int gemm_tn<T>(Ctx*, int64_t, int64_t, int64_t, double, const T*, int64_t, const T*, int64_t, double, T*, int64_t, int, double*);
INST(double)
INST(float)

}  // namespace rlb
