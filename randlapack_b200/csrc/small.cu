// Small / streaming helper kernels of the sketch-and-factor path:
//   potrf_upper      lapack::potrf(Upper) on the k x k Gram matrix           (rl_orth.hh:81)
//   trtri_upper      explicit inverse of the Cholesky factor, so that the trsm of rl_orth.hh:95 becomes a
//                    tall GEMM on the tensor pipe
//   jacobi_svd       lapack::gesdd(SomeVec) of the tall n x k matrix B^T     (rl_rsvd.hh:146) by one-sided
//                    Jacobi (Hestenes) with a round-robin ordering on a cooperative grid
//   sumsq            lapack::lange(Fro)                                      (rl_qb.hh:168,221)
// All are latency- or HBM-bound and far off the critical path of the tall GEMMs.
#include "common.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <cmath>

namespace cg = cooperative_groups;

namespace rlb {

// ------------------------------------------------------------------------------------------------
// potrf (upper): A = R^T R, R overwrites the upper triangle; strictly-lower part untouched.
// One CTA; row j of R is produced from rows < j (left-looking), all threads work on row j's columns.
// info = 0, or j+1 for the first non-positive (or NaN) pivot, like LAPACK.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024) potrf_upper_kernel(int k, T* __restrict__ A, int lda, int* __restrict__ info) {
    __shared__ double s_piv;
    __shared__ int s_fail;
    if (threadIdx.x == 0) s_fail = 0;
    __syncthreads();
    for (int j = 0; j < k; ++j) {
        // diagonal: ajj = A[j][j] - sum_{i<j} R[i][j]^2  (column j is contiguous)
        if (threadIdx.x < 32) {
            double s = 0.0;
            for (int i = threadIdx.x; i < j; i += 32) { double r = (double)A[i + (size_t)j * lda]; s += r * r; }
            s = warp_sum(s);
            if (threadIdx.x == 0) {
                double ajj = (double)A[j + (size_t)j * lda] - s;
                if (!(ajj > 0.0)) { s_fail = j + 1; s_piv = 1.0; }
                else { s_piv = sqrt(ajj); A[j + (size_t)j * lda] = (T)s_piv; }
            }
        }
        __syncthreads();
        if (s_fail) break;
        const double piv = s_piv;
        // row j of R: R[j][c] = (A[j][c] - sum_{i<j} R[i][j] R[i][c]) / R[j][j], one thread per column c > j
        for (int c = j + 1 + threadIdx.x; c < k; c += blockDim.x) {
            const T* cj = A + (size_t)j * lda;
            const T* cc = A + (size_t)c * lda;
            double s = 0.0;
            for (int i = 0; i < j; ++i) s += (double)cj[i] * (double)cc[i];
            A[j + (size_t)c * lda] = (T)(((double)cc[j] - s) / piv);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *info = s_fail;
}

// Rinv (k x k, upper, ld k; strictly-lower zeroed) = inverse of upper-triangular R.  One thread per column.
template <typename T>
__global__ void trtri_upper_kernel(int k, const T* __restrict__ R, int ldr, T* __restrict__ Rinv) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k) return;
    T* x = Rinv + (size_t)c * k;
    for (int i = c + 1; i < k; ++i) x[i] = (T)0;
    x[c] = (T)(1.0 / (double)R[c + (size_t)c * ldr]);
    for (int i = c - 1; i >= 0; --i) {
        double s = 0.0;
        for (int l = i + 1; l <= c; ++l) s += (double)R[i + (size_t)l * ldr] * (double)x[l];
        x[i] = (T)(-s / (double)R[i + (size_t)i * ldr]);
    }
}

// ------------------------------------------------------------------------------------------------
// sum of squares (two-stage, deterministic): partial[b] per block, then final in one block
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const T* __restrict__ A, int64_t m, int64_t n, int64_t lda,
                                                            double* __restrict__ partial) {
    double s = 0.0;
    const int64_t total = m * n;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = e / m, i = e - j * m;
        double v = (double)A[i + j * lda];
        s += v * v;
    }
    __shared__ double sh[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        partial[blockIdx.x] = t;
    }
}
__global__ void sum_final_kernel(const double* __restrict__ partial, int nb, double* __restrict__ out) {
    // single warp, fixed order
    double s = 0.0;
    for (int i = threadIdx.x; i < nb; i += 32) s += partial[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) *out = s;
}

template <typename T>
int sumsq(Ctx* ctx, const T* A, int64_t m, int64_t n, int64_t lda, double* partial_ws, double* out_dev) {
    const int64_t total = m * n;
    int nb = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)ctx->num_sms * 8));
    LaunchScope ls(ctx, RLB200_TIMER_SMALL, 2);
    sumsq_partial_kernel<T><<<nb, 256, 0, ctx->stream>>>(A, m, n, lda, partial_ws);
    sum_final_kernel<<<1, 32, 0, ctx->stream>>>(partial_ws, nb, out_dev);
    RLB_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}
int sumsq_ws_doubles(Ctx* ctx) { return ctx->num_sms * 8; }

template <typename T>
int potrf_upper(Ctx* ctx, int k, T* A, int lda, int* info_dev) {
    LaunchScope ls(ctx, RLB200_TIMER_SMALL);
    potrf_upper_kernel<T><<<1, 1024, 0, ctx->stream>>>(k, A, lda, info_dev);
    RLB_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}
template <typename T>
int trtri_upper(Ctx* ctx, int k, const T* R, int ldr, T* Rinv) {
    LaunchScope ls(ctx, RLB200_TIMER_SMALL);
    trtri_upper_kernel<T><<<(k + 63) / 64, 64, 0, ctx->stream>>>(k, R, ldr, Rinv);
    RLB_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// One-sided Jacobi SVD of a tall n x k matrix B (ld n):  B <- B J with orthogonal columns, W <- J.
// Round-robin (circle method) ordering: k' = k rounded up to even, k'-1 rounds per sweep, k'/2 disjoint
// column pairs per round, one CTA per pair (grid-stride), grid.sync() between rounds.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) jacobi_svd_kernel(int n, int k, T* __restrict__ B, int ldb, T* __restrict__ W, int max_sweeps,
                                                         double tol, int* __restrict__ rot_count, int* __restrict__ sweeps_done) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[3][8];
    __shared__ double s_c, s_s;
    __shared__ int s_rot;
    const int kp = (k + 1) & ~1;
    const int npairs = kp / 2;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        for (int round = 0; round < kp - 1; ++round) {
            for (int pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
                int p, q;
                if (pr == 0) { p = kp - 1; q = round; }
                else { p = (round + pr) % (kp - 1); q = (round - pr + (kp - 1)) % (kp - 1); }
                if (p > q) { int t = p; p = q; q = t; }
                if (q >= k) continue;   // dummy column of the odd case (uniform across the CTA)
                T* bp = B + (size_t)p * ldb;
                T* bq = B + (size_t)q * ldb;
                double a = 0.0, b = 0.0, g = 0.0;
                for (int i = tid; i < n; i += 256) {
                    double x = (double)bp[i], y = (double)bq[i];
                    a += x * x; b += y * y; g += x * y;
                }
                a = warp_sum(a); b = warp_sum(b); g = warp_sum(g);
                if (lane == 0) { sh[0][wid] = a; sh[1][wid] = b; sh[2][wid] = g; }
                __syncthreads();
                if (tid == 0) {
                    a = b = g = 0.0;
                    for (int w = 0; w < 8; ++w) { a += sh[0][w]; b += sh[1][w]; g += sh[2][w]; }
                    int rot = 0;
                    double c = 1.0, s = 0.0;
                    if (fabs(g) > tol * sqrt(a * b) && g != 0.0) {
                        double zeta = (b - a) / (2.0 * g);
                        double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                        c = 1.0 / sqrt(1.0 + t * t);
                        s = c * t;
                        rot = 1;
                    }
                    s_c = c; s_s = s; s_rot = rot;
                    if (rot) atomicAdd(rot_count, 1);
                }
                __syncthreads();
                if (s_rot) {
                    const double c = s_c, s = s_s;
                    for (int i = tid; i < n; i += 256) {
                        double x = (double)bp[i], y = (double)bq[i];
                        bp[i] = (T)(c * x - s * y);
                        bq[i] = (T)(s * x + c * y);
                    }
                    T* wp = W + (size_t)p * k;
                    T* wq = W + (size_t)q * k;
                    for (int i = tid; i < k; i += 256) {
                        double x = (double)wp[i], y = (double)wq[i];
                        wp[i] = (T)(c * x - s * y);
                        wq[i] = (T)(s * x + c * y);
                    }
                }
                __syncthreads();
            }
            grid.sync();
        }
        // converged when a whole sweep applied no rotation
        const int rots = *reinterpret_cast<volatile int*>(rot_count);
        grid.sync();
        if (blockIdx.x == 0 && tid == 0) { *rot_count = 0; *sweeps_done = sweep + 1; }
        grid.sync();
        if (rots == 0) break;
    }
}

// sigma_j = ||B_j||, order = indices sorted by descending sigma (stable), single CTA (k <= 4096)
template <typename T>
__global__ void __launch_bounds__(256) svd_norms_kernel(int n, int k, const T* __restrict__ B, int ldb, double* __restrict__ sig) {
    const int j = blockIdx.x;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) { double x = (double)B[i + (size_t)j * ldb]; s += x * x; }
    __shared__ double sh[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < 8; ++w) t += sh[w]; sig[j] = sqrt(t); }
}
__global__ void svd_rank_kernel(int k, const double* __restrict__ sig, int* __restrict__ dest) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    const double sj = sig[j];
    int r = 0;
    for (int l = 0; l < k; ++l) { double sl = sig[l]; r += (sl > sj) || (sl == sj && l < j); }
    dest[j] = r;   // column j goes to position r
}
// out[:, dest[j]] = in[:, j] * scale_j   (scale = 1/sigma for the left vectors, 1 for W)
template <typename T>
__global__ void __launch_bounds__(256) svd_permute_kernel(int rows, int k, const T* __restrict__ in, int ldin, T* __restrict__ out,
                                                          int ldout, const int* __restrict__ dest, const double* __restrict__ sig,
                                                          int normalise, T* __restrict__ S_out) {
    const int j = blockIdx.x;
    const int d = dest[j];
    const double sg = sig[j];
    const double sc = normalise ? (sg > 0.0 ? 1.0 / sg : 0.0) : 1.0;
    for (int i = threadIdx.x; i < rows; i += 256) out[i + (size_t)d * ldout] = (T)((double)in[i + (size_t)j * ldin] * sc);
    if (S_out && threadIdx.x == 0) S_out[d] = (T)sg;
}
template <typename T>
__global__ void set_identity_kernel(int k, T* __restrict__ W) {
    const int64_t total = (int64_t)k * k;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
        W[e] = (e % k == e / k) ? (T)1 : (T)0;
}

// B (n x k) -> left singular vectors, S -> singular values (descending), W (k x k) -> right singular vectors.
// ws: scratch of at least svd_ws_bytes(n, k).
size_t svd_ws_bytes(int64_t n, int64_t k, size_t elem) {
    return ws_round((size_t)n * k * elem) + ws_round((size_t)k * k * elem) + ws_round(k * sizeof(double)) + ws_round(k * sizeof(int)) + 512;
}

template <typename T>
int svd_tall(Ctx* ctx, int64_t n, int64_t k, T* B, int64_t ldb, T* S, T* W, void* ws, int* sweeps_out) {
    RLB_REQUIRE(ctx, n >= k && k >= 1 && n < (1ll << 31) && k <= 16384);
    WsCarver cv(ws);
    T* Bt = cv.take<T>((size_t)n * k);
    T* Wt = cv.take<T>((size_t)k * k);
    double* sig = cv.take<double>(k);
    int* dest = cv.take<int>(k);
    int* flags = cv.take<int>(2);
    RLB_CUDA_OK(ctx, cudaMemsetAsync(flags, 0, 2 * sizeof(int), ctx->stream));
    {
        LaunchScope ls(ctx, RLB200_TIMER_SMALL, 6);
        set_identity_kernel<T><<<std::min<int64_t>((k * k + 255) / 256, 1024), 256, 0, ctx->stream>>>((int)k, Wt);
        int occ = 0;
        RLB_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, jacobi_svd_kernel<T>, 256, 0));
        int npairs = (int)((k + 1) / 2);
        int grid = std::max(1, std::min(npairs, occ * ctx->num_sms));
        int n_ = (int)n, k_ = (int)k, ldb_ = (int)ldb, max_sweeps = 60;
        double tol = (sizeof(T) == 8 ? 2.220446049250313e-16 : 1.1920929e-07) * std::sqrt((double)n);
        int* rot = flags; int* swp = flags + 1;
        void* args[] = {&n_, &k_, &B, &ldb_, &Wt, &max_sweeps, &tol, &rot, &swp};
        RLB_CUDA_OK(ctx, cudaLaunchCooperativeKernel((void*)jacobi_svd_kernel<T>, dim3(grid), dim3(256), args, 0, ctx->stream));
        svd_norms_kernel<T><<<(unsigned)k, 256, 0, ctx->stream>>>((int)n, (int)k, B, (int)ldb, sig);
        svd_rank_kernel<<<(unsigned)((k + 127) / 128), 128, 0, ctx->stream>>>((int)k, sig, dest);
        RLB_CUDA_OK(ctx, cudaMemcpy2DAsync(Bt, n * sizeof(T), B, ldb * sizeof(T), n * sizeof(T), k, cudaMemcpyDeviceToDevice, ctx->stream));
        svd_permute_kernel<T><<<(unsigned)k, 256, 0, ctx->stream>>>((int)n, (int)k, Bt, (int)n, B, (int)ldb, dest, sig, 1, S);
        svd_permute_kernel<T><<<(unsigned)k, 256, 0, ctx->stream>>>((int)k, (int)k, Wt, (int)k, W, (int)k, dest, sig, 0, (T*)nullptr);
        RLB_CUDA_OK(ctx, cudaGetLastError());
    }
    if (sweeps_out) {
        RLB_CUDA_OK(ctx, cudaMemcpyAsync(ctx->hbox, flags + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        *sweeps_out = *static_cast<int*>(ctx->hbox);
    }
    return 0;
}

#define INST(T)                                                                   \
    template int sumsq<T>(Ctx*, const T*, int64_t, int64_t, int64_t, double*, double*); \
    template int potrf_upper<T>(Ctx*, int, T*, int, int*);                        \
    template int trtri_upper<T>(Ctx*, int, const T*, int, T*);                    \
    template int svd_tall<T>(Ctx*, int64_t, int64_t, T*, int64_t, T*, T*, void*, int*);
INST(double)
INST(float)

}  // namespace rlb
