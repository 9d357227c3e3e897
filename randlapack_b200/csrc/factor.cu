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
#include <cooperative_groups.h>
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

// The same step for TALL panels (d - j in the tens of thousands): one CTA per trailing column instead of one warp, so that a column's d - j
// entries are covered by 256 threads (a warp walking 32768 rows is 1024 dependent iterations: 150 us per step measured through BQRRP / hqrrp
// at m = 32768).  The sum of squares of the updated column is accumulated in the same pass, so the norm recomputation costs nothing extra.
template <typename T, bool PIVOT>
__global__ void __launch_bounds__(256) qr_apply_block_kernel(int64_t d, int64_t n, int64_t j, T* __restrict__ A, int64_t lda, double* __restrict__ vn1,
                                                             double* __restrict__ vn2, const QrcpStep* __restrict__ step, double tol3z) {
    __shared__ double s_red[8];
    __shared__ double s_bc;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double tau = step->tau;
    const T* v = A + j * lda;
    auto block_sum = [&](double x) -> double {
        x = warp_sum(x);
        __syncthreads();                       // s_red / s_bc of the previous reduction have been read by everyone
        if (lane == 0) s_red[wid] = x;
        __syncthreads();
        if (tid == 0) s_bc = ((s_red[0] + s_red[1]) + (s_red[2] + s_red[3])) + ((s_red[4] + s_red[5]) + (s_red[6] + s_red[7]));
        __syncthreads();
        return s_bc;
    };
    for (int64_t c = j + 1 + blockIdx.x; c < n; c += gridDim.x) {
        T* a = A + c * lda;
        double ajc = (double)a[j];
        double ss = 0.0;
        if (tau != 0.0) {
            double w = 0.0;
            for (int64_t i = j + 1 + tid; i < d; i += 256) w = fma((double)v[i], (double)a[i], w);
            w = block_sum(w) + ajc;
            const double tw = tau * w;
            for (int64_t i = j + 1 + tid; i < d; i += 256) {
                const T nv = (T)((double)a[i] - tw * (double)v[i]);
                a[i] = nv;
                ss = fma((double)nv, (double)nv, ss);
            }
            ajc = (double)(T)(ajc - tw);
            if (tid == 0) a[j] = (T)ajc;
        } else if (PIVOT) {
            for (int64_t i = j + 1 + tid; i < d; i += 256) { const double x = (double)a[i]; ss = fma(x, x, ss); }
        }
        if (PIVOT) {
            ss = block_sum(ss);
            const double n1 = vn1[c];
            if (n1 != 0.0) {                   // LAPACK dlaqp2 partial-norm downdate
                double temp = fabs(ajc) / n1;
                temp = fmax(0.0, (1.0 + temp) * (1.0 - temp));
                const double r = n1 / vn2[c];
                const double temp2 = temp * r * r;
                __syncthreads();               // every thread has read vn1 / vn2 of this column
                if (tid == 0) {
                    if (temp2 <= tol3z) { const double nv = (j + 1 < d) ? sqrt(ss) : 0.0; vn1[c] = nv; vn2[c] = nv; }
                    else vn1[c] = n1 * sqrt(temp);
                }
            }
        }
    }
}

template <typename T>
__global__ void iota_i64_kernel(int64_t n, int64_t* p, int64_t base) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = base + i;
}

// ------------------------------------------------------------------------------------------------
// The same factorisation in ONE cooperative launch (one grid-wide barrier per column instead of two kernel launches).
//   * columns are never swapped physically: a column keeps its storage and carries a logical position (col2pos); idamax ties are broken by
//     position, which reproduces LAPACK's order after its swaps; the columns are gathered into pivot order once at the end;
//   * nothing is serial: every CTA reduces the per-CTA pivot candidates of the previous step, reads alpha and the exact sum of squares of
//     the pivot column below the diagonal (accumulated for every column for free while it was last updated) and derives dlarfg's
//     (beta, tau, scale) redundantly; the scaled reflector is staged in shared memory; a warp owns a column: dot, update, sum of squares
//     of the updated column, dlaqp2 norm downdate and the next pivot candidate in one pass pair;
//   * the pivot column itself is scaled / gets beta one step later by its owner (other CTAs still read its raw entries during the step).
// ------------------------------------------------------------------------------------------------
struct QrCand { double v; long long pos; long long col; long long pad; };

__device__ __forceinline__ bool qr_better(double v, long long pos, double bv, long long bp) { return v > bv || (v == bv && pos < bp); }

// A column is owned by a group of GS threads: one warp for short columns (d <= 1024), four warps beyond.
constexpr int kQrEpt = 32;            // column entries a thread keeps in registers: one pass over memory for up to GS * 32 live rows

template <int GS>
__device__ __forceinline__ void qr_gbar(int grp) {
    if constexpr (GS == 32) __syncwarp();
    else asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
}
// sum over the warps of a group; every thread of the group gets the total.  `slot` alternates so that consecutive reductions do not collide.
template <int GS>
__device__ __forceinline__ double qr_group_sum(double v, double* red /* [2][4] of this group */, int slot, int grp, int wig, int lane) {
    v = warp_sum(v);
    if constexpr (GS == 32) return v;
    if (lane == 0) red[slot * 4 + wig] = v;
    qr_gbar<GS>(grp);
    return (red[slot * 4 + 0] + red[slot * 4 + 1]) + (red[slot * 4 + 2] + red[slot * 4 + 3]);
}

template <typename T, bool PIVOT, int GS>
__global__ void __launch_bounds__(512) qr_coop_kernel(int64_t d, int64_t n, int64_t kmin, T* __restrict__ A, int64_t lda, double* __restrict__ vn1,
                                                      double* __restrict__ vn2, double* __restrict__ ss2, long long* __restrict__ col2pos,
                                                      QrCand* __restrict__ cand, T* __restrict__ tau_out, double safmin, double tol3z) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double sv[];                        // scaled reflector, rows j+1 .. d-1
    constexpr int kQrGroups = 512 / GS;
    __shared__ double s_bv[kQrGroups];
    __shared__ long long s_bp[kQrGroups], s_bc[kQrGroups];
    __shared__ double s_red[kQrGroups][12];
    __shared__ double s_sc[4];                            // scal, tau, beta
    __shared__ long long s_piv[2];                        // pivot column, its position
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = tid / GS, wig = (tid % GS) >> 5, gt = tid % GS;     // group, warp in group, thread in group
    const int64_t G = gridDim.x, cta = blockIdx.x;
    const int64_t gg = cta * kQrGroups + grp, ng = G * kQrGroups;      // columns c == gg (mod ng) belong to this group
    double* red = s_red[grp];

    auto publish = [&](double bv, long long bp, long long bc, int64_t slot) {
        if (gt == 0) { s_bv[grp] = bv; s_bp[grp] = bp; s_bc[grp] = bc; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kQrGroups; ++w) if (s_bc[w] >= 0 && (bc < 0 || qr_better(s_bv[w], s_bp[w], bv, bp))) { bv = s_bv[w]; bp = s_bp[w]; bc = s_bc[w]; }
            QrCand c; c.v = bv; c.pos = bp; c.col = bc; c.pad = 0;
            cand[slot * G + cta] = c;
        }
    };

    // prologue: column norms, sum of squares below row 0, logical positions, first candidates
    {
        double bv = -1.0; long long bp = n, bc = -1;
        for (int64_t c = gg; c < n; c += ng) {
            const T* a = A + c * lda;
            double s = 0.0;
            for (int64_t i = 1 + gt; i < d; i += GS) { const double x = (double)a[i]; s = fma(x, x, s); }
            s = qr_group_sum<GS>(s, red, 0, grp, wig, lane);
            const double a0 = (double)a[0];
            const double nv = sqrt(fma(a0, a0, s));
            if (gt == 0) { ss2[c] = s; col2pos[c] = c; if (PIVOT) { vn1[c] = nv; vn2[c] = nv; } }
            if (PIVOT) { const double key = (nv == nv) ? nv : -0.5; if (bc < 0 || qr_better(key, c, bv, bp)) { bv = key; bp = c; bc = c; } }
            qr_gbar<GS>(grp);      // red[] is free again
        }
        if (PIVOT) publish(bv, bp, bc, 0);
    }
    grid.sync();

    long long prev_pc = -1;
    double prev_scal = 1.0, prev_beta = 0.0;
    bool prev_scaled = false;
    for (int64_t j = 0; j < kmin; ++j) {
        // ---- pivot (every CTA, redundantly)
        if (warp == 0) {
            long long pc = j, pp = j;
            if (PIVOT) {
                double bv = -2.0; long long bp = n; pc = -1;
                for (int64_t g = lane; g < G; g += 32) {
                    const QrCand* q = cand + (j & 1) * G + g;
                    const double v = __ldcg(&q->v); const long long ps = __ldcg(&q->pos), cl = __ldcg(&q->col);
                    if (cl >= 0 && (pc < 0 || qr_better(v, ps, bv, bp))) { bv = v; bp = ps; pc = cl; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const long long op = __shfl_xor_sync(0xffffffffu, bp, o), oc = __shfl_xor_sync(0xffffffffu, pc, o);
                    if (oc >= 0 && (pc < 0 || qr_better(ov, op, bv, bp))) { bv = ov; bp = op; pc = oc; }
                }
                pp = bp;
            }
            if (lane == 0) {
                s_piv[0] = pc; s_piv[1] = pp;
                // dlarfg(d - j, x[j], x[j+1:]) from alpha and the exact sum of squares below the diagonal
                const double alpha = (double)__ldcg(A + j + pc * lda);
                const double xnorm = (j + 1 < d) ? sqrt(__ldcg(ss2 + pc)) : 0.0;
                double tau = 0.0, scal = 1.0, beta = alpha;
                if (xnorm != 0.0 && j + 1 < d) {
                    beta = -copysign(hypot(alpha, xnorm), alpha);
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
                s_sc[0] = scal; s_sc[1] = tau; s_sc[2] = beta;
                if (cta == 0) tau_out[j] = (T)tau;
            }
        }
        __syncthreads();
        const long long pc = s_piv[0], pp = s_piv[1];
        const double scal = s_sc[0], tau_full = s_sc[1], beta = s_sc[2];
        const double tau = (double)(T)tau_full;
        if (tau_full != 0.0) {
            const T* x = A + pc * lda;
            for (int64_t i = j + 1 + tid; i < d; i += blockDim.x) sv[i] = (double)(T)((double)__ldcg(x + i) * scal);
        }
        __syncthreads();
        // ---- apply H_j to the owned live columns: a group of 4 warps per column, the column's live rows in registers
        const int64_t len = d - (j + 1);
        const bool one_pass = len <= (int64_t)GS * kQrEpt;
        double bv = -1.0; long long bp = n, bc = -1;
        for (int64_t c = gg; c < n; c += ng) {
            T* a = A + c * lda;
            long long pos = col2pos[c];
            if (c == pc) { qr_gbar<GS>(grp); if (gt == 0) col2pos[c] = j; continue; }
            if (pos < j) {
                if (c == prev_pc) {                       // last step's pivot column: nobody reads it any more
                    if (prev_scaled) for (int64_t i = j + gt; i < d; i += GS) a[i] = (T)((double)a[i] * prev_scal);
                    if (gt == 0) a[j - 1] = (T)prev_beta;
                }
                continue;
            }
            qr_gbar<GS>(grp);      // everyone has read col2pos[c] before it may change
            if (pos == j) { pos = pp; if (gt == 0) col2pos[c] = pp; }          // the displaced column takes the pivot's old position
            double ajc = (double)a[j];
            double s = 0.0, a_next = 0.0;
            T ar[kQrEpt];
            if (one_pass) {
#pragma unroll
                for (int e = 0; e < kQrEpt; ++e) { const int64_t i = j + 1 + gt + GS * e; ar[e] = (i < d) ? a[i] : (T)0; }
            }
            if (tau_full != 0.0) {
                double w = 0.0;
                if (one_pass) {
#pragma unroll
                    for (int e = 0; e < kQrEpt; ++e) { const int64_t i = j + 1 + gt + GS * e; if (i < d) w = fma(sv[i], (double)ar[e], w); }
                } else {
                    for (int64_t i = j + 1 + gt; i < d; i += GS) w = fma(sv[i], (double)a[i], w);
                }
                w = qr_group_sum<GS>(w, red, 0, grp, wig, lane) + ajc;
                const double tw = tau * w;
                if (one_pass) {
#pragma unroll
                    for (int e = 0; e < kQrEpt; ++e) {
                        const int64_t i = j + 1 + gt + GS * e;
                        if (i < d) {
                            const T nvT = (T)((double)ar[e] - tw * sv[i]);
                            a[i] = nvT;
                            const double xn = (double)nvT;
                            if (i == j + 1) a_next = xn; else s = fma(xn, xn, s);
                        }
                    }
                } else {
                    for (int64_t i = j + 1 + gt; i < d; i += GS) {
                        const T nvT = (T)((double)a[i] - tw * sv[i]);
                        a[i] = nvT;
                        const double xn = (double)nvT;
                        if (i == j + 1) a_next = xn; else s = fma(xn, xn, s);
                    }
                }
                ajc = (double)(T)(ajc - tw);
                if (gt == 0) a[j] = (T)ajc;
            } else {
                if (one_pass) {
#pragma unroll
                    for (int e = 0; e < kQrEpt; ++e) {
                        const int64_t i = j + 1 + gt + GS * e;
                        if (i < d) { const double xn = (double)ar[e]; if (i == j + 1) a_next = xn; else s = fma(xn, xn, s); }
                    }
                } else {
                    for (int64_t i = j + 1 + gt; i < d; i += GS) { const double xn = (double)a[i]; if (i == j + 1) a_next = xn; else s = fma(xn, xn, s); }
                }
            }
            if constexpr (GS == 32) {
                a_next = __shfl_sync(0xffffffffu, a_next, 0);     // row j + 1 is held by thread 0 of the group
                s = warp_sum(s);
            } else {
                if (gt == 0) red[8] = a_next;
                s = qr_group_sum<GS>(s, red, 1, grp, wig, lane);
                a_next = red[8];
            }
            if (gt == 0) ss2[c] = s;                      // rows >= j + 2: what dlarfg needs if this column is the next pivot
            if (PIVOT) {
                double n1 = vn1[c];
                if (n1 != 0.0) {                          // LAPACK dlaqp2 partial-norm downdate
                    double temp = fabs(ajc) / n1;
                    temp = fmax(0.0, (1.0 + temp) * (1.0 - temp));
                    const double r = n1 / vn2[c];
                    const double temp2 = temp * r * r;
                    if (temp2 <= tol3z) {
                        n1 = (j + 1 < d) ? sqrt(fma(a_next, a_next, s)) : 0.0;
                        qr_gbar<GS>(grp);  // all threads of the group have read vn1 / vn2
                        if (gt == 0) { vn1[c] = n1; vn2[c] = n1; }
                    } else {
                        n1 = n1 * sqrt(temp);
                        qr_gbar<GS>(grp);
                        if (gt == 0) vn1[c] = n1;
                    }
                }
                const double key = (n1 == n1) ? n1 : -0.5;
                if (bc < 0 || qr_better(key, pos, bv, bp)) { bv = key; bp = pos; bc = c; }
            }
        }
        prev_pc = pc; prev_scal = scal; prev_beta = beta; prev_scaled = tau_full != 0.0;
        if (PIVOT) publish(bv, bp, bc, (j + 1) & 1);
        grid.sync();
    }
    // the last pivot column
    for (int64_t c = gg; c < n; c += ng) {
        if (c != prev_pc) continue;
        T* a = A + c * lda;
        if (prev_scaled) for (int64_t i = kmin + gt; i < d; i += GS) a[i] = (T)((double)a[i] * prev_scal);
        if (gt == 0) a[kmin - 1] = (T)prev_beta;
    }
}

// pivots and the gather into pivot order: position q receives the column stored at pos2col[q]
__global__ void qr_pos2col_kernel(int64_t n, const long long* __restrict__ col2pos, long long* __restrict__ pos2col, int64_t* __restrict__ jpvt) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
        const long long q = col2pos[c];
        pos2col[q] = c;
        jpvt[q] = c + 1;
    }
}
template <typename T>
__global__ void __launch_bounds__(256) qr_gather_kernel(int64_t d, int64_t n, const T* __restrict__ src, T* __restrict__ A, int64_t lda,
                                                        const long long* __restrict__ pos2col) {
    for (int64_t q = blockIdx.x; q < n; q += gridDim.x) {
        const T* s = src + (int64_t)pos2col[q] * d;
        T* a = A + q * lda;
        for (int64_t i = threadIdx.x; i < d; i += blockDim.x) a[i] = s[i];
    }
}

size_t qrcp_ws_bytes(int64_t n) { return ws_round(sizeof(double) * n) * 2 + ws_round(sizeof(QrcpStep)); }

// geqp3 (pivot = true; jpvt_dev receives 1-based GEQP3-style pivots, all columns free) or geqrf (pivot = false; jpvt_dev unused)
// of the d x n matrix A (lda); tau_dev has min(d, n) entries.  R ends up in the upper triangle, reflectors below it.
// stages >= 0: only the first min(stages, d, n) Householder steps are taken (the num_stages of NoFLA_QRPmod_WY_unb_var4, rl_hqrrp.hh:556-575:
// jpvt then holds exactly the column swaps of those steps).  tol3z_in > 0 replaces sqrt(eps of T) in the norm-downdate recompute rule
// (rl_hqrrp.hh:376-378 takes sqrt(dlamch('E')) for every T).
template <typename T>
int qr_small(Ctx* ctx, bool pivot, int64_t d, int64_t n, T* A, int64_t lda, int64_t* jpvt_dev, T* tau_dev, void* ws, int64_t stages, double tol3z_in) {
    RLB_REQUIRE(ctx, d >= 0 && n >= 0 && lda >= std::max<int64_t>(d, 1));
    const int64_t kmin = stages >= 0 ? std::min(stages, std::min(d, n)) : std::min(d, n);
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
    const double tol3z = tol3z_in > 0.0 ? tol3z_in : std::sqrt(eps);
    // one cooperative launch (qr_coop_kernel) whenever the reflector fits shared memory and the grid can be co-resident
    // (measured, tools/bench_qrcp.py: 4096 x 2048: 91.7 -> 41.2 ms; with more than ~8 columns per group and step the two-launch form, whose apply
    //  kernel spreads the columns over 8 CTAs per SM, is faster: 256 x 65536: 19.8 vs 35.9 ms)
    const bool wide_groups = d > 1024;
    const int groups_per_cta = wide_groups ? 4 : 16;
    if (d <= 16384 && n <= (int64_t)8 * groups_per_cta * ctx->num_sms && getenv("RLB200_QR_NOCOOP") == nullptr) {
        const size_t smem = sizeof(double) * (size_t)d;
        void* kern = wide_groups ? (pivot ? (void*)qr_coop_kernel<T, true, 128> : (void*)qr_coop_kernel<T, false, 128>)
                                 : (pivot ? (void*)qr_coop_kernel<T, true, 32> : (void*)qr_coop_kernel<T, false, 32>);
        RLB_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        RLB_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 512, smem));
        if (occ > 0) {
            ArenaScope as(ctx);
            const int64_t G = std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms, (n + groups_per_cta - 1) / groups_per_cta));
            double* ss2 = as.take<double>((size_t)n); if (!ss2) return RLB200_ERR_ALLOC;
            long long* col2pos = as.take<long long>((size_t)n); if (!col2pos) return RLB200_ERR_ALLOC;
            long long* pos2col = as.take<long long>((size_t)n); if (!pos2col) return RLB200_ERR_ALLOC;
            QrCand* cand = as.take<QrCand>((size_t)2 * G); if (!cand) return RLB200_ERR_ALLOC;
            T* tmp = nullptr;
            if (pivot) { tmp = as.take<T>((size_t)d * n); if (!tmp) return RLB200_ERR_ALLOC; }
            LaunchScope ls(ctx, RLB200_TIMER_FACTOR, pivot ? 3 : 1);
            int64_t d_ = d, n_ = n, k_ = kmin, lda_ = lda;
            double safmin_ = safmin, tol3z_ = tol3z;
            void* args[] = {&d_, &n_, &k_, &A, &lda_, &vn1, &vn2, &ss2, &col2pos, &cand, &tau_dev, &safmin_, &tol3z_};
            RLB_CUDA_OK(ctx, cudaLaunchCooperativeKernel(kern, dim3((unsigned)G), dim3(512), args, smem, ctx->stream));
            if (pivot) {
                RLB_CUDA_OK(ctx, cudaMemcpy2DAsync(tmp, d * sizeof(T), A, lda * sizeof(T), d * sizeof(T), n, cudaMemcpyDeviceToDevice, ctx->stream));
                qr_pos2col_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 1024), 256, 0, ctx->stream>>>(n, col2pos, pos2col, jpvt_dev);
                qr_gather_kernel<T><<<(unsigned)std::min<int64_t>(n, (int64_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(d, n, tmp, A, lda, pos2col);
            }
            RLB_CUDA_OK(ctx, cudaGetLastError());
            return 0;      // (the scratch is reused in stream order)
        }
    }
    LaunchScope ls(ctx, RLB200_TIMER_FACTOR, (int)(2 * kmin + 2));
    if (pivot) {
        iota_i64_kernel<T><<<(unsigned)std::min<int64_t>((n + 255) / 256, 1024), 256, 0, ctx->stream>>>(n, jpvt_dev, 1);
        colnorm_kernel<T><<<(unsigned)std::min<int64_t>((n + 7) / 8, (int64_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(d, n, A, lda, vn1, vn2);
    }
    const int head_threads = d >= 2048 ? 1024 : (d >= 512 ? 512 : 256);
    static const bool warp_apply_env = getenv("RLB200_QR_WARP_APPLY") != nullptr;      // timing switch: the one-warp-per-column apply everywhere
    for (int64_t j = 0; j < kmin; ++j) {
        if (pivot) qr_head_kernel<T, true><<<1, head_threads, 0, ctx->stream>>>(d, n, j, A, lda, vn1, vn2, jpvt_dev, tau_dev, step, safmin);
        else       qr_head_kernel<T, false><<<1, head_threads, 0, ctx->stream>>>(d, n, j, A, lda, vn1, vn2, jpvt_dev, tau_dev, step, safmin);
        const int64_t rem = n - j - 1;
        if (rem > 0 && d - j >= 4096 && !warp_apply_env) {
            const unsigned nbk = (unsigned)std::min<int64_t>(rem, (int64_t)ctx->num_sms * 8);
            if (pivot) qr_apply_block_kernel<T, true><<<nbk, 256, 0, ctx->stream>>>(d, n, j, A, lda, vn1, vn2, step, tol3z);
            else       qr_apply_block_kernel<T, false><<<nbk, 256, 0, ctx->stream>>>(d, n, j, A, lda, vn1, vn2, step, tol3z);
        } else if (rem > 0) {
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
    template int qr_small<T>(Ctx*, bool, int64_t, int64_t, T*, int64_t, int64_t*, T*, void*, int64_t, double);
INST(double)
INST(float)

}  // namespace rlb
