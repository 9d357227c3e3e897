// PLUL stabiliser (RandLAPACK/comps/rl_orth.hh:211-230): getrf with partial pivoting on the m x n iterate,
// keep the unit-lower-trapezoidal L (util::get_L, rl_util.hh:101-114) and apply lapack::laswp(n, A, m, 1, n, ipiv, +1).
//
// Pivot rule = LAPACK's: at step j the pivot is the FIRST row of maximal |a_ij| (idamax) among rows i >= j of the
// updated column; the sub-diagonal is scaled by the reciprocal of the pivot; a zero pivot column is skipped.
// Right-looking blocked form: panels of 32 columns are factored with the unblocked BLAS-2 steps (pivot search over the whole
// column, interchange across all n columns, scale + rank-1 update inside the panel only); the rest of the row block is then
// solved with the panel's unit-lower triangle and the trailing matrix gets one rank-32 update on the tall tensor-pipe GEMM.
// HBM traffic drops from 8*m*n^2 bytes to ~8*m*n^2/32; everything stays on the device (pivot indices live in device memory).
#include "drivers.cuh"
#include <algorithm>
#include <cstdlib>
#include <cooperative_groups.h>

namespace rlb {

struct PivCand {
    double v;
    long long i;
};

__device__ __forceinline__ PivCand better(PivCand a, PivCand b) {
    // larger |value| wins; ties go to the smaller row index (idamax returns the first maximum)
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}

template <typename T>
__global__ void __launch_bounds__(256) lu_argmax_partial(const T* __restrict__ A, int64_t m, int64_t lda, int j, PivCand* __restrict__ part) {
    PivCand best{-1.0, (long long)m};
    const T* col = A + (int64_t)j * lda;
    for (int64_t i = j + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        PivCand c{fabs((double)col[i]), (long long)i};
        best = better(best, c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        PivCand c{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
        best = better(best, c);
    }
    __shared__ PivCand sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) best = better(best, sh[w]);
        part[blockIdx.x] = best;
    }
}

// final reduction + row interchange j <-> p over all n columns; records ipiv[j] = p (0-based) and pivot validity
template <typename T>
__global__ void __launch_bounds__(256) lu_pivot_swap(T* __restrict__ A, int64_t lda, int n, int j, const PivCand* __restrict__ part, int nparts,
                                                     long long* __restrict__ ipiv, int* __restrict__ nonzero) {
    __shared__ PivCand sbest;
    if (threadIdx.x < 32) {
        PivCand best{-1.0, (long long)1 << 62};
        for (int q = threadIdx.x; q < nparts; q += 32) best = better(best, part[q]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            PivCand c{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
            best = better(best, c);
        }
        if (threadIdx.x == 0) { sbest = best; ipiv[j] = best.i; nonzero[j] = best.v != 0.0; }
    }
    __syncthreads();
    const long long p = sbest.i;
    if (sbest.v != 0.0 && p != j) {
        for (int c = threadIdx.x; c < n; c += blockDim.x) {
            T* col = A + (int64_t)c * lda;
            T t = col[j]; col[j] = col[p]; col[p] = t;
        }
    }
}

// l_i = a_ij * (1 / a_jj) for i > j, then a_ic -= l_i * a_jc for c > j   (dscal by the reciprocal + dger)
template <typename T>
__global__ void __launch_bounds__(256) lu_scale_update(T* __restrict__ A, int64_t m, int64_t lda, int n, int j, const int* __restrict__ nonzero) {
    __shared__ double urow[1024];
    const int ncols = n - j - 1;
    T* colj = A + (int64_t)j * lda;
    const bool nz = nonzero[j] != 0;
    const T rinv = nz ? (T)1 / colj[j] : (T)1;
    for (int c0 = 0; c0 < max(ncols, 1); c0 += 1024) {
        const int nc = min(1024, ncols - c0);
        __syncthreads();
        for (int c = threadIdx.x; c < nc; c += blockDim.x) urow[c] = (double)A[j + (int64_t)(j + 1 + c0 + c) * lda];
        __syncthreads();
        for (int64_t i = j + 1 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
            T l = colj[i];
            if (c0 == 0 && nz) { l = l * rinv; colj[i] = l; }
            else if (nz) { /* already scaled in the first chunk */ }
            for (int c = 0; c < nc; ++c) {
                T* a = A + i + (int64_t)(j + 1 + c0 + c) * lda;
                *a = (T)((double)*a - (double)l * urow[c]);
            }
        }
        if (ncols <= 0) break;
    }
}

// get_L (zero strictly-upper part, unit diagonal) followed by laswp(n, A, lda, 1, n, ipiv, +1):
// for i = 0..kmin-1 in order: swap rows i and ipiv[i]
template <typename T>
__global__ void __launch_bounds__(256) lu_make_L(T* __restrict__ A, int64_t m, int64_t lda, int n) {
    const int64_t total = (int64_t)n * min((int64_t)n, m);   // only the top min(m,n) rows hold upper-triangle entries
    const int64_t rows = min((int64_t)n, m);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e % rows, c = e / rows;
        if (i < c) A[i + c * lda] = (T)0;
        else if (i == c) A[i + c * lda] = (T)1;
    }
}
template <typename T>
__global__ void __launch_bounds__(1024) lu_laswp_forward(T* __restrict__ A, int64_t lda, int n, int kmin, const long long* __restrict__ ipiv) {
    for (int i = 0; i < kmin; ++i) {
        const long long p = ipiv[i];
        if (p != i) {
            for (int c = threadIdx.x; c < n; c += blockDim.x) {
                T* col = A + (int64_t)c * lda;
                T t = col[i]; col[i] = col[p]; col[p] = t;
            }
        }
        __syncthreads();
    }
}

// A12 (nb x ncols) <- L11^-1 A12, L11 = unit lower triangle of the nb x nb block at L (nb <= 32).  One thread per column.
template <typename T>
__global__ void __launch_bounds__(128) lu_trsm_unit_lower(const T* __restrict__ L, int64_t lda, int nb, T* __restrict__ A12, int ncols) {
    __shared__ double sl[32][33];
    for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) sl[e % nb][e / nb] = (double)L[(e % nb) + (int64_t)(e / nb) * lda];
    __syncthreads();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    T* col = A12 + (int64_t)c * lda;
    double x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = i < nb ? (double)col[i] : 0.0;
#pragma unroll
    for (int i = 1; i < 32; ++i) {
        if (i < nb) {
            double s = x[i];
#pragma unroll
            for (int k = 0; k < i; ++k) s -= sl[i][k] * x[k];
            x[i] = s;
        }
    }
#pragma unroll
    for (int i = 1; i < 32; ++i) if (i < nb) col[i] = (T)x[i];
}

// One 32-column panel [jb, jend) of the blocked LU in ONE cooperative launch: per column a grid-wide pivot reduction, the row
// interchange across all n columns, and the scale + rank-1 update of the panel's remaining columns, which also yields the pivot
// candidates of the next column (two grid syncs per column instead of three launches).
template <typename T>
__global__ void __launch_bounds__(256) lu_panel_coop_kernel(T* __restrict__ A, int64_t m, int64_t lda, int n, int jb, int jend,
                                                            PivCand* __restrict__ part, long long* __restrict__ ipiv, int* __restrict__ nonzero) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ PivCand sh[8];
    __shared__ PivCand sbest;
    __shared__ double urow[32];
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
    auto block_publish = [&](PivCand best) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            PivCand c{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
            best = better(best, c);
        }
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) best = better(best, sh[w]);
            part[blockIdx.x] = best;
        }
        __syncthreads();
    };
    {   // pivot candidates of the first column of the panel
        PivCand best{-1.0, (long long)m};
        const T* col = A + (int64_t)jb * lda;
        for (int64_t i = jb + gtid; i < m; i += gsz) best = better(best, PivCand{fabs((double)col[i]), (long long)i});
        block_publish(best);
    }
    for (int j = jb; j < jend; ++j) {
        grid.sync();
        // every CTA reduces the candidates (same order everywhere); CTA 0 records the pivot and interchanges rows j and p
        if (threadIdx.x < 32) {
            PivCand best{-1.0, (long long)1 << 62};
            for (int q = threadIdx.x; q < (int)gridDim.x; q += 32) best = better(best, part[q]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                PivCand c{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
                best = better(best, c);
            }
            if (threadIdx.x == 0) sbest = best;
        }
        __syncthreads();
        const long long p = sbest.i;
        const bool nz = sbest.v != 0.0;
        if (blockIdx.x == 0) {
            if (threadIdx.x == 0) { ipiv[j] = p; nonzero[j] = nz ? 1 : 0; }
            if (nz && p != j) {
                for (int c = threadIdx.x; c < n; c += blockDim.x) {
                    T* col = A + (int64_t)c * lda;
                    const T t = col[j]; col[j] = col[p]; col[p] = t;
                }
            }
        }
        grid.sync();
        // scale the pivot column, rank-1 update of the panel's remaining columns; candidates of column j + 1 on the fly
        const int nc = jend - j - 1;
        T* colj = A + (int64_t)j * lda;
        if (threadIdx.x < nc) urow[threadIdx.x] = (double)A[j + (int64_t)(j + 1 + threadIdx.x) * lda];
        __syncthreads();
        const T rinv = nz ? (T)1 / colj[j] : (T)1;
        PivCand best{-1.0, (long long)m};
        for (int64_t i = j + 1 + gtid; i < m; i += gsz) {
            T l = colj[i];
            if (nz) { l = l * rinv; colj[i] = l; }
            for (int c = 0; c < nc; ++c) {
                T* a = A + i + (int64_t)(j + 1 + c) * lda;
                const T v = (T)((double)*a - (double)l * urow[c]);
                *a = v;
                if (c == 0) best = better(best, PivCand{fabs((double)v), (long long)i});
            }
        }
        if (j + 1 < jend) block_publish(best);
    }
}

size_t plul_ws_bytes(Ctx* ctx, int64_t n) {
    return ws_round(sizeof(PivCand) * (size_t)ctx->num_sms * 8) + ws_round(sizeof(long long) * n) + ws_round(sizeof(int) * n);
}

// lapack::getrf(m, n, A, lda, ipiv) without host round trips: ipiv_dev[j] = 0-based pivot row of step j, nonzero_dev[j] = pivot != 0.
// ws: plul_ws_bytes(ctx, n) scratch (the partial-argmax candidates live at its start).
template <typename T>
int getrf_nopiv_out(Ctx* ctx, int64_t m, int64_t n, T* A, int64_t lda, PivCand* part, long long* ipiv, int* nonzero) {
    const int kmin = (int)std::min<int64_t>(m, n);
    constexpr int NB = 32;
    for (int jb = 0; jb < kmin; jb += NB) {
        const int jend = std::min(jb + NB, kmin);
        const int64_t prow = m - jb;
        int occ = 0;
        RLB_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lu_panel_coop_kernel<T>, 256, 0));
        if (occ > 0 && prow >= 4096 && getenv("RLB200_LU_NOCOOP") == nullptr) {
            // one cooperative launch per panel (grid = all co-resident CTAs that have rows to own)
            const int gridc = (int)std::max<int64_t>(1, std::min<int64_t>((prow + 255) / 256, std::min<int64_t>((int64_t)occ * ctx->num_sms, (int64_t)ctx->num_sms * 8)));
            LaunchScope ls(ctx, RLB200_TIMER_SMALL);
            int n_ = (int)n, jb_ = jb, jend_ = jend;
            void* args[] = {&A, &m, &lda, &n_, &jb_, &jend_, &part, &ipiv, &nonzero};
            RLB_CUDA_OK(ctx, cudaLaunchCooperativeKernel((void*)lu_panel_coop_kernel<T>, dim3(gridc), dim3(256), args, 0, ctx->stream));
        } else {
            LaunchScope ls(ctx, RLB200_TIMER_SMALL, 3 * (jend - jb));
            for (int j = jb; j < jend; ++j) {
                const int64_t rows = m - j;
                const int nb = (int)std::max<int64_t>(1, std::min<int64_t>((rows + 255) / 256, (int64_t)ctx->num_sms * 8));
                lu_argmax_partial<T><<<nb, 256, 0, ctx->stream>>>(A, m, lda, j, part);
                lu_pivot_swap<T><<<1, 256, 0, ctx->stream>>>(A, lda, (int)n, j, part, nb, ipiv, nonzero);
                if (rows > 1) {
                    // scale the pivot column, rank-1 update of the panel's remaining columns only
                    const int ub = (int)std::max<int64_t>(1, std::min<int64_t>((rows - 1 + 255) / 256, (int64_t)ctx->num_sms * 8));
                    lu_scale_update<T><<<ub, 256, 0, ctx->stream>>>(A, m, lda, jend, j, nonzero);
                }
            }
            RLB_CUDA_OK(ctx, cudaGetLastError());
        }
        if (jend < n) {
            const int rest = (int)n - jend;
            {
                LaunchScope ls(ctx, RLB200_TIMER_SMALL);
                lu_trsm_unit_lower<T><<<(rest + 127) / 128, 128, 0, ctx->stream>>>(A + jb + (int64_t)jb * lda, lda, jend - jb, A + jb + (int64_t)jend * lda, rest);
                RLB_CUDA_OK(ctx, cudaGetLastError());
            }
            if (m > jend)
                RLB_CHECK(gemm_nn<T>(ctx, m - jend, rest, jend - jb, -1.0, A + jend + (int64_t)jb * lda, lda, A + jb + (int64_t)jend * lda, lda, 1.0,
                                     A + jend + (int64_t)jend * lda, lda));
        }
    }
    return 0;
}

template <typename T>
int plul(Ctx* ctx, int64_t m, int64_t n, T* A, int64_t lda, void* ws) {
    RLB_REQUIRE(ctx, m >= 0 && n >= 0 && n < (1 << 30));
    if (m == 0 || n == 0) return 0;
    WsCarver cv(ws);
    PivCand* part = cv.take<PivCand>((size_t)ctx->num_sms * 8);
    long long* ipiv = cv.take<long long>(n);
    int* nonzero = cv.take<int>(n);
    const int kmin = (int)std::min<int64_t>(m, n);
    RLB_CHECK(getrf_nopiv_out<T>(ctx, m, n, A, lda, part, ipiv, nonzero));
    LaunchScope ls(ctx, RLB200_TIMER_SMALL, 2);
    lu_make_L<T><<<(unsigned)std::min<int64_t>((n * std::min<int64_t>(n, m) + 255) / 256, 4096), 256, 0, ctx->stream>>>(A, m, lda, (int)n);
    lu_laswp_forward<T><<<1, 1024, 0, ctx->stream>>>(A, lda, (int)n, kmin, ipiv);
    RLB_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

// getrf with the pivot vector returned to the HOST (0-based rows), for BQRRP's LU-based QRCP (rl_bqrrp.hh:341-356)
template <typename T>
int getrf_pivots(Ctx* ctx, int64_t m, int64_t n, T* A, int64_t lda, void* ws, std::vector<int64_t>& ipiv_host) {
    RLB_REQUIRE(ctx, m >= 0 && n >= 0 && n < (1 << 30));
    const int kmin = (int)std::min<int64_t>(m, n);
    ipiv_host.assign((size_t)kmin, 0);
    if (kmin == 0) return 0;
    WsCarver cv(ws);
    PivCand* part = cv.take<PivCand>((size_t)ctx->num_sms * 8);
    long long* ipiv = cv.take<long long>(n);
    int* nonzero = cv.take<int>(n);
    RLB_CHECK(getrf_nopiv_out<T>(ctx, m, n, A, lda, part, ipiv, nonzero));
    static_assert(sizeof(long long) == sizeof(int64_t), "");
    RLB_CUDA_OK(ctx, cudaMemcpyAsync(ipiv_host.data(), ipiv, sizeof(int64_t) * kmin, cudaMemcpyDeviceToHost, ctx->stream));
    RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

template int plul<double>(Ctx*, int64_t, int64_t, double*, int64_t, void*);
template int plul<float>(Ctx*, int64_t, int64_t, float*, int64_t, void*);
template int getrf_pivots<double>(Ctx*, int64_t, int64_t, double*, int64_t, void*, std::vector<int64_t>&);
template int getrf_pivots<float>(Ctx*, int64_t, int64_t, float*, int64_t, void*, std::vector<int64_t>&);

}  // namespace rlb
