// Factorisation building blocks of CQRRPT / BQRRP on device.
//
//   col_permute        util::col_swap = lapack::lapmt(forward)            RandLAPACK/misc/rl_util.hh:151-165 (used rl_cqrrpt.hh:292, rl_bqrrp.hh:359,365)
//   copy_tri / set_tri lapack::lacpy(Upper), util::get_U, laset          rl_cqrrpt.hh:284, rl_util.hh:119-131
//   potrf_blocked      lapack::potrf(Upper) for k beyond one CTA          rl_cqrrpt.hh:311
//   trsm_right_upper   blas::trsm(Right, Upper, NoTrans, NonUnit)         rl_cqrrpt.hh:306,342, rl_bqrrp.hh:443,450,606 — blocked, in place,
//                      off-diagonal blocks through the tall DMMA GEMM, diagonal blocks through the triangular in-place GEMM
//   qrcp               lapack::geqp3 (LAPACK dgeqp3/dlaqp2 pivot rule: first column of maximal downdated partial norm, Drmac
//                      downdate with the tol3z = sqrt(eps) recompute rule, dlarfg reflectors) rl_cqrrpt.hh:247, rl_bqrrp.hh:336;
//                      also the norm machinery of rl_hqrrp.hh:336-461
//   geqrf_unblocked    lapack::geqrf of the small d x n sketch            rl_bqrrp.hh:356
#include "drivers.cuh"
#include <algorithm>
#include <cmath>
#include <limits>

namespace rlb {

// ------------------------------------------------------------------------------------------------
// column permutation, in place: new column i = old column perm[i] (0-based), cycle following.
// HBM-bound: every element of a moved column is read once and written once (2*sizeof(T)*m*n bytes at most).
// Threads own rows, so a thread walks all cycles for its rows without any synchronisation; a warp touches 32 consecutive rows of
// one column per access (coalesced).
// ------------------------------------------------------------------------------------------------
template <typename T, typename V>
__global__ void __launch_bounds__(256) col_permute_kernel(T* __restrict__ A, int64_t lda_v, int64_t m_v, const int* __restrict__ nodes,
                                                          const int* __restrict__ starts, int ncycles) {
    V* Av = reinterpret_cast<V*>(A);
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m_v; r += (int64_t)gridDim.x * blockDim.x) {
        for (int c = 0; c < ncycles; ++c) {
            const int s = starts[c], e = starts[c + 1];
            const V first = Av[r + (int64_t)nodes[s] * lda_v];
            V cur = Av[r + (int64_t)nodes[s + 1] * lda_v];
            for (int t = s; t < e - 1; ++t) {
                // prefetch the next source before storing the current one
                V nxt = cur;
                if (t + 2 < e) nxt = Av[r + (int64_t)nodes[t + 2] * lda_v];
                Av[r + (int64_t)nodes[t] * lda_v] = cur;
                cur = nxt;
            }
            Av[r + (int64_t)nodes[e - 1] * lda_v] = first;
        }
    }
}

// perm_host: 0-based, length n, a permutation of 0..n-1 (new col i = old col perm[i])
template <typename T>
int col_permute(Ctx* ctx, int64_t m, int64_t n, T* A, int64_t lda, const int64_t* perm_host) {
    if (m == 0 || n == 0) return 0;
    std::vector<int> nodes, starts;
    std::vector<char> seen((size_t)n, 0);
    for (int64_t i = 0; i < n; ++i) {
        if (seen[i] || perm_host[i] == i) { seen[i] = 1; continue; }
        starts.push_back((int)nodes.size());
        int64_t j = i;
        while (!seen[j]) { seen[j] = 1; nodes.push_back((int)j); j = perm_host[j]; RLB_REQUIRE(ctx, j >= 0 && j < n); }
    }
    if (starts.empty()) return 0;
    starts.push_back((int)nodes.size());
    const int ncycles = (int)starts.size() - 1;
    ArenaScope as(ctx);
    int* d_nodes = as.take<int>(nodes.size()); if (!d_nodes) return RLB200_ERR_ALLOC;
    int* d_starts = as.take<int>(starts.size()); if (!d_starts) return RLB200_ERR_ALLOC;
    RLB_CUDA_OK(ctx, cudaMemcpyAsync(d_nodes, nodes.data(), nodes.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    RLB_CUDA_OK(ctx, cudaMemcpyAsync(d_starts, starts.data(), starts.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));   // the host vectors die at return
    LaunchScope ls(ctx, RLB200_TIMER_FACTOR);
    constexpr int E = 16 / sizeof(T);
    const bool vec = (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (lda % E == 0) && (m % E == 0);
    if (vec) {
        const int64_t mv = m / E;
        const int nb = (int)std::min<int64_t>((mv + 255) / 256, (int64_t)ctx->num_sms * 16);
        col_permute_kernel<T, double2><<<nb, 256, 0, ctx->stream>>>(A, lda / E, mv, d_nodes, d_starts, ncycles);
    } else {
        const int nb = (int)std::min<int64_t>((m + 255) / 256, (int64_t)ctx->num_sms * 16);
        col_permute_kernel<T, T><<<nb, 256, 0, ctx->stream>>>(A, lda, m, d_nodes, d_starts, ncycles);
    }
    RLB_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// triangle helpers.  mode 0: dst upper triangle (incl. diagonal) <- src; 1: zero the strictly lower part of dst (get_U);
// 2: dst <- upper triangle of src, strictly lower part zeroed (clean copy)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) tri_kernel(int mode, int64_t rows, int64_t cols, const T* __restrict__ src, int64_t lds, T* __restrict__ dst,
                                                  int64_t ldd) {
    const int64_t total = rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e % rows, j = e / rows;
        if (mode == 0) { if (i <= j) dst[i + j * ldd] = src[i + j * lds]; }
        else if (mode == 1) { if (i > j) dst[i + j * ldd] = (T)0; }
        else dst[i + j * ldd] = (i <= j) ? src[i + j * lds] : (T)0;
    }
}
template <typename T>
int tri_op(Ctx* ctx, int mode, int64_t rows, int64_t cols, const T* src, int64_t lds, T* dst, int64_t ldd) {
    if (rows <= 0 || cols <= 0) return 0;
    LaunchScope ls(ctx, RLB200_TIMER_SMALL);
    const int nb = (int)std::min<int64_t>((rows * cols + 255) / 256, (int64_t)ctx->num_sms * 8);
    tri_kernel<T><<<nb, 256, 0, ctx->stream>>>(mode, rows, cols, src, lds, dst, ldd);
    RLB_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// blocked Cholesky (upper), right-looking with NB = 128: diagonal block by the one-CTA kernel, block row by
// R12 = R11^{-T} A12 (explicit 128 x 128 inverse + tensor GEMM), trailing update A22 -= R12^T R12 (tensor GEMM, upper tiles).
// info_host = 0 or the 1-based index of the first non-positive pivot (LAPACK semantics).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void add_info_offset_kernel(int* info, int offset, int* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { if (*out == 0 && *info != 0) *out = *info + offset; }
}

// dst[i] = A[i, i] (save) or A[i, i] = src[i] for i >= from (restore)
template <typename T>
__global__ void __launch_bounds__(256) diag_save_restore_kernel(T* __restrict__ A, int64_t lda, int64_t k, int64_t from, T* __restrict__ buf, int restore) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= k || i < from) return;
    if (restore) A[i + i * lda] = buf[i]; else buf[i] = A[i + i * lda];
}

template <typename T>
int potrf_blocked(Ctx* ctx, int64_t k, T* A, int64_t lda, int* info_host) {
    constexpr int NB = 128;
    *info_host = 0;
    if (k == 0) return 0;
    ArenaScope as(ctx);
    int* info_dev = as.take<int>(1); if (!info_dev) return RLB200_ERR_ALLOC;
    if (k <= 128) {      // (the single-CTA kernel is latency-bound: 2.1 ms at k = 256 under ncu; two 128-blocks + GEMM updates take a third)
        RLB_CHECK(potrf_upper<T>(ctx, (int)k, A, (int)lda, info_dev));
    } else {
        T* inv = as.take<T>(NB * NB); if (!inv) return RLB200_ERR_ALLOC;
        T* tmp = as.take<T>((size_t)NB * k); if (!tmp) return RLB200_ERR_ALLOC;
        // This factorization is right-looking (the trailing matrix carries Schur complements); LAPACK's blocked dpotrf, which the reference
        // calls, is left-looking and on failure leaves the diagonal blocks after the failing one untouched.  CQRRPT's a-posteriori rank
        // estimate walks the whole diagonal after a failure (rl_cqrrpt.hh:319-331), so the original diagonal is kept and restored there.
        T* diag0 = as.take<T>((size_t)k); if (!diag0) return RLB200_ERR_ALLOC;
        diag_save_restore_kernel<T><<<(unsigned)((k + 255) / 256), 256, 0, ctx->stream>>>(A, lda, k, 0, diag0, 0);
        for (int64_t j = 0; j < k; j += NB) {
            const int64_t jb = std::min<int64_t>(NB, k - j), rem = k - j - jb;
            T* Ajj = A + j + j * lda;
            RLB_CHECK(potrf_upper<T>(ctx, (int)jb, Ajj, (int)lda, info_dev));
            int info = 0;
            RLB_CUDA_OK(ctx, cudaMemcpyAsync(ctx->hbox, info_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
            info = *static_cast<int*>(ctx->hbox);
            if (info != 0) {
                *info_host = info + (int)j;
                if (j + jb < k) diag_save_restore_kernel<T><<<(unsigned)((k + 255) / 256), 256, 0, ctx->stream>>>(A, lda, k, j + jb, diag0, 1);
                RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));      // diag0 is scratch of this scope
                return 0;
            }
            if (rem > 0) {
                T* A12 = A + j + (j + jb) * lda;
                RLB_CHECK(trtri_upper<T>(ctx, (int)jb, Ajj, (int)lda, inv));
                // tmp(jb x rem) = inv^T * A12
                RLB_CHECK(gemm_tn<T>(ctx, jb, jb, rem, 1.0, inv, jb, A12, lda, 0.0, tmp, jb, 0));
                RLB_CUDA_OK(ctx, cudaMemcpy2DAsync(A12, lda * sizeof(T), tmp, jb * sizeof(T), jb * sizeof(T), rem, cudaMemcpyDeviceToDevice, ctx->stream));
                // A22 -= R12^T R12 (upper tiles only)
                RLB_CHECK(gemm_tn<T>(ctx, jb, rem, rem, -1.0, A12, lda, A12, lda, 1.0, A + (j + jb) + (j + jb) * lda, lda, 1));
            }
        }
        return 0;
    }
    RLB_CUDA_OK(ctx, cudaMemcpyAsync(ctx->hbox, info_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    *info_host = *static_cast<int*>(ctx->hbox);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// X <- X * R^{-1}, X m x k (tall, in place), R k x k upper triangular (ldr).  Column blocks of 64, left to right:
//   X_j <- (X_j - X[:, 0:j0] * R[0:j0, j]) * R_jj^{-1}
// The off-diagonal products run through the 128x64x32 two-CTA-per-SM tile (the fastest DMMA configuration measured); the
// diagonal 64 x 64 solves are triangular-skipping in-place products with the explicit inverse of the diagonal block.
// (One 256-wide in-place tile instead costs 2.2x the time at k = 256: 1 CTA/SM, 204 registers — profiles/launches_full_r1.csv.)
// ------------------------------------------------------------------------------------------------
template <typename T>
int trsm_right_upper(Ctx* ctx, int64_t m, int64_t k, const T* R, int64_t ldr, T* X, int64_t ldx) {
    constexpr int NB = 64;
    if (m == 0 || k == 0) return 0;
    ArenaScope as(ctx);
    T* inv = as.take<T>(NB * NB); if (!inv) return RLB200_ERR_ALLOC;
    for (int64_t j0 = 0; j0 < k; j0 += NB) {
        const int64_t jb = std::min<int64_t>(NB, k - j0);
        T* Xj = X + j0 * ldx;
        if (j0 > 0) RLB_CHECK(gemm_nn<T>(ctx, m, jb, j0, -1.0, X, ldx, R + j0 * ldr, ldr, 1.0, Xj, ldx));
        RLB_CHECK(trtri_upper<T>(ctx, (int)jb, R + j0 + j0 * ldr, (int)ldr, inv));
        RLB_CHECK(gemm_nn_inplace<T>(ctx, m, jb, jb, 1.0, Xj, ldx, inv, jb, /*b_upper_tri=*/true));
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// QRCP / QR of a small d x n matrix (the sketch), unblocked Householder (LAPACK dlaqp2 / dgeqr2 arithmetic):
// two launches per column — (1) one CTA: pivot search over the downdated norms, column swap, dlarfg;
// (2) one warp per trailing column: w = v^T a_c, a_c -= tau w v, norm downdate (recompute when cancellation is detected).
// The matrix (<= a few tens of MB) lives in L2; the loop is launch/latency-bound (~2 x 3 us per column).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colnorm_kernel(int64_t d, int64_t n, const T* __restrict__ A, int64_t lda, double* __restrict__ vn1,
                                                      double* __restrict__ vn2) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int64_t c = (int64_t)blockIdx.x * 8 + wib; c < n; c += (int64_t)gridDim.x * 8) {
        const T* a = A + c * lda;
        double s = 0.0;
        for (int64_t i = lane; i < d; i += 32) { const double v = (double)a[i]; s = fma(v, v, s); }
        s = warp_sum(s);
        if (lane == 0) { const double nv = sqrt(s); vn1[c] = nv; if (vn2) vn2[c] = nv; }
    }
}

struct QrcpStep {          // per-step scalars produced by the head kernel and consumed by the apply kernel
    double tau;
    int valid;
};

// head of step j: (pivot &) reflector.  PIVOT = false gives plain geqr2.
template <typename T, bool PIVOT>
__global__ void __launch_bounds__(1024) qr_head_kernel(int64_t d, int64_t n, int64_t j, T* __restrict__ A, int64_t lda, double* __restrict__ vn1,
                                                       double* __restrict__ vn2, int64_t* __restrict__ jpvt, T* __restrict__ tau_out,
                                                       QrcpStep* __restrict__ step, double safmin) {
    __shared__ double sv[32];
    __shared__ long long si[32];
    __shared__ long long s_p;
    __shared__ double s_x[2];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    if (PIVOT) {
        // idamax over vn1[j..n): first index of the maximum
        double bv = -1.0; long long bi = n;
        for (int64_t c = j + tid; c < n; c += blockDim.x) { const double v = vn1[c]; if (v > bv) { bv = v; bi = c; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o); const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { sv[wid] = bv; si[wid] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < nw; ++w) if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
            if (bi >= n) bi = j;    // all-NaN guard
            s_p = bi;
            if (bi != j) {
                const int64_t t = jpvt[bi]; jpvt[bi] = jpvt[j]; jpvt[j] = t;
                vn1[bi] = vn1[j]; vn2[bi] = vn2[j];
            }
        }
        __syncthreads();
        const int64_t p = s_p;
        if (p != j) {
            T* a = A + j * lda; T* b = A + p * lda;
            for (int64_t i = tid; i < d; i += blockDim.x) { const T t = a[i]; a[i] = b[i]; b[i] = t; }
        }
        __syncthreads();
    }
    // dlarfg(d - j, A[j][j], A[j+1:, j])
    T* col = A + j * lda;
    double ss = 0.0;
    for (int64_t i = j + 1 + tid; i < d; i += blockDim.x) { const double v = (double)col[i]; ss = fma(v, v, ss); }
    ss = warp_sum(ss);
    if (lane == 0) sv[wid] = ss;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < nw; ++w) t += sv[w];
        const double xnorm = sqrt(t);
        const double alpha = (double)col[j];
        double tau = 0.0, scal = 1.0, beta = alpha;
        if (xnorm != 0.0 && j + 1 < d) {
            beta = -copysign(hypot(alpha, xnorm), alpha);
            // (the safmin rescaling loop of dlarfg matters only for |beta| < ~1e-292 / 1e-30; handled by scaling in fp64 for float
            //  inputs, and by a single rescale for doubles)
            if (fabs(beta) < safmin) {
                const double rs = 1.0 / safmin;
                const double a2 = alpha * rs, x2 = xnorm * rs;
                const double b2 = -copysign(hypot(a2, x2), a2);
                tau = (b2 - a2) / b2;
                scal = rs / (a2 - b2);
                beta = b2 * safmin;
            } else {
                tau = (beta - alpha) / beta;
                scal = 1.0 / (alpha - beta);
            }
        }
        s_x[0] = scal; s_x[1] = tau;
        col[j] = (T)beta;
        tau_out[j] = (T)tau;
        step->tau = (double)(T)tau;
        step->valid = 1;
    }
    __syncthreads();
    const double scal = s_x[0];
    if (s_x[1] != 0.0)
        for (int64_t i = j + 1 + tid; i < d; i += blockDim.x) col[i] = (T)((double)col[i] * scal);
}

// apply H_j = I - tau v v^T (v = [1; A[j+1:, j]]) to the trailing columns; one warp per column.
template <typename T, bool PIVOT>
__global__ void __launch_bounds__(256) qr_apply_kernel(int64_t d, int64_t n, int64_t j, T* __restrict__ A, int64_t lda, double* __restrict__ vn1,
                                                       double* __restrict__ vn2, const QrcpStep* __restrict__ step, double tol3z) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const double tau = step->tau;
    const T* v = A + j * lda;
    for (int64_t c = j + 1 + (int64_t)blockIdx.x * 8 + wib; c < n; c += (int64_t)gridDim.x * 8) {
        T* a = A + c * lda;
        double ajc = (double)a[j];
        if (tau != 0.0) {
            double w = 0.0;
            for (int64_t i = j + 1 + lane; i < d; i += 32) w = fma((double)v[i], (double)a[i], w);
            w = warp_sum(w) + ajc;
            const double tw = tau * w;
            for (int64_t i = j + 1 + lane; i < d; i += 32) a[i] = (T)((double)a[i] - tw * (double)v[i]);
            ajc = (double)(T)(ajc - tw);
            if (lane == 0) a[j] = (T)ajc;
        }
        if (PIVOT) {
            // LAPACK dlaqp2 partial-norm downdate
            const double n1 = vn1[c];
            if (n1 != 0.0) {
                double temp = fabs(ajc) / n1;
                temp = fmax(0.0, (1.0 + temp) * (1.0 - temp));
                const double r = n1 / vn2[c];
                const double temp2 = temp * r * r;
                if (temp2 <= tol3z) {
                    double s = 0.0;
                    if (j + 1 < d) {
                        __syncwarp();
                        for (int64_t i = j + 1 + lane; i < d; i += 32) { const double x = (double)a[i]; s = fma(x, x, s); }
                        s = warp_sum(s);
                    }
                    if (lane == 0) { const double nv = sqrt(s); vn1[c] = nv; vn2[c] = nv; }
                } else if (lane == 0) {
                    vn1[c] = n1 * sqrt(temp);
                }
            }
        }
    }
}

template <typename T>
__global__ void iota_i64_kernel(int64_t n, int64_t* p, int64_t base) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = base + i;
}

size_t qrcp_ws_bytes(int64_t n) { return ws_round(sizeof(double) * n) * 2 + ws_round(sizeof(QrcpStep)); }

// geqp3 (pivot = true; jpvt_dev receives 1-based GEQP3-style pivots, all columns free) or geqrf (pivot = false; jpvt_dev unused)
// of the d x n matrix A (lda); tau_dev has min(d, n) entries.  R ends up in the upper triangle, reflectors below it.
template <typename T>
int qr_small(Ctx* ctx, bool pivot, int64_t d, int64_t n, T* A, int64_t lda, int64_t* jpvt_dev, T* tau_dev, void* ws) {
    RLB_REQUIRE(ctx, d >= 0 && n >= 0 && lda >= std::max<int64_t>(d, 1));
    const int64_t kmin = std::min(d, n);
    if (kmin == 0) {
        if (pivot && n > 0) { iota_i64_kernel<T><<<(unsigned)std::min<int64_t>((n + 255) / 256, 1024), 256, 0, ctx->stream>>>(n, jpvt_dev, 1); }
        return 0;
    }
    WsCarver cv(ws);
    double* vn1 = cv.take<double>(n);
    double* vn2 = cv.take<double>(n);
    QrcpStep* step = cv.take<QrcpStep>(1);
    const double eps = sizeof(T) == 8 ? 1.1102230246251565e-16 : 5.9604644775390625e-08;       // lamch('Epsilon')
    const double safmin = (sizeof(T) == 8 ? 2.2250738585072014e-308 : 1.1754943508222875e-38) / eps;
    const double tol3z = std::sqrt(eps);
    LaunchScope ls(ctx, RLB200_TIMER_FACTOR, (int)(2 * kmin + 2));
    if (pivot) {
        iota_i64_kernel<T><<<(unsigned)std::min<int64_t>((n + 255) / 256, 1024), 256, 0, ctx->stream>>>(n, jpvt_dev, 1);
        colnorm_kernel<T><<<(unsigned)std::min<int64_t>((n + 7) / 8, (int64_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(d, n, A, lda, vn1, vn2);
    }
    const int head_threads = d >= 2048 ? 1024 : (d >= 512 ? 512 : 256);
    for (int64_t j = 0; j < kmin; ++j) {
        if (pivot) qr_head_kernel<T, true><<<1, head_threads, 0, ctx->stream>>>(d, n, j, A, lda, vn1, vn2, jpvt_dev, tau_dev, step, safmin);
        else       qr_head_kernel<T, false><<<1, head_threads, 0, ctx->stream>>>(d, n, j, A, lda, vn1, vn2, jpvt_dev, tau_dev, step, safmin);
        const int64_t rem = n - j - 1;
        if (rem > 0) {
            const unsigned nb = (unsigned)std::min<int64_t>((rem + 7) / 8, (int64_t)ctx->num_sms * 8);
            if (pivot) qr_apply_kernel<T, true><<<nb, 256, 0, ctx->stream>>>(d, n, j, A, lda, vn1, vn2, step, tol3z);
            else       qr_apply_kernel<T, false><<<nb, 256, 0, ctx->stream>>>(d, n, j, A, lda, vn1, vn2, step, tol3z);
        }
    }
    RLB_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

#define INST(T)                                                                                              \
    template int col_permute<T>(Ctx*, int64_t, int64_t, T*, int64_t, const int64_t*);                        \
    template int tri_op<T>(Ctx*, int, int64_t, int64_t, const T*, int64_t, T*, int64_t);                     \
    template int potrf_blocked<T>(Ctx*, int64_t, T*, int64_t, int*);                                         \
    template int trsm_right_upper<T>(Ctx*, int64_t, int64_t, const T*, int64_t, T*, int64_t);                \
    template int qr_small<T>(Ctx*, bool, int64_t, int64_t, T*, int64_t, int64_t*, T*, void*);
INST(double)
INST(float)

}  // namespace rlb
