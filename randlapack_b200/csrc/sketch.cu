// RandBLAS sketch-apply on device.
//
//   sparse (SASO / count-sketch)   SparseDist + SparseSkOp + fill_sparse(_unpacked)  RandBLAS/RandBLAS/sparse_skops.hh:55-142, 167-312, 568-704
//                                  sample_indices_iid_uniform                        RandBLAS/RandBLAS/util.hh:520-542
//                                  lskges -> left_spmm -> apply_csr_jik_p11          RandBLAS/RandBLAS/skge.hh:538-571, sparse_data/csr_spmm_impl.hh:115-153
//   dense (Gaussian / uniform)     lskge3 / rskge3                                   RandBLAS/RandBLAS/skge.hh:155-203, 308-356
//
// Sparse left sketch  B(d x n) = alpha * S(d x m) * A(m x n) + beta * B,  S wide, Axis::Short, vec_nnz non-zeros (+-1) per column.
// HBM-bound: A is streamed exactly once (sizeof(T)*m*n bytes), the operator costs 2 bytes per non-zero.  No atomics
// (shared-memory float atomics are CAS loops on sm_100a and L2 reductions run at ~1.3 cycles/lane): instead
//   1. saso_plan_kernel regenerates the operator from the Philox state, one CTA per chunk of R consecutive columns of S (= rows of A),
//      and sorts the chunk's non-zeros by (sketch row, source row): a per-chunk CSR (offsets per sketch row + packed local row/sign).
//   2. saso_apply_kernel: a CTA owns CW columns of A and all d sketch rows; every thread owns RPT sketch rows and keeps its
//      RPT x CW accumulators in REGISTERS for its whole row range; chunks of A are staged through shared memory with a cp.async
//      double buffer; a thread walks the CSR lists of its rows and adds the staged values.  Summation order is fixed
//      (source rows ascending), so the result is run-to-run deterministic.
//   3. partial sums of the row splits are combined in a fixed order.
#include "drivers.cuh"
#include "philox.cuh"
#include <algorithm>
#include <cstdlib>

namespace rlb {

constexpr int kSasoMaxNnz = 64;
constexpr int kSasoThreads = 512;

// ---- one column of a wide Axis::Short SparseSkOp: the draws of repeated_fisher_yates (sparse_skops.hh:55-142) ----------------------
// rows[t], neg[t] in draw order, then insertion-sorted by row (sparse_skops.hh:641-668; rows are distinct).
__device__ void saso_column(const Ctr128& seed, uint32_t k0, uint32_t k1, int64_t j, int nnz, int64_t d, uint32_t* rows, uint8_t* neg) {
    uint32_t lp[2 * kSasoMaxNnz], lv[2 * kSasoMaxNnz];   // log of assignments to the (virtually identity) index array
    int nlog = 0;
    auto get = [&](uint32_t x) {
        for (int i = nlog - 1; i >= 0; --i) if (lp[i] == x) return lv[i];
        return x;
    };
    for (int t = 0; t < nnz; ++t) {
        uint32_t rv[4];
        philox4x32_10(ctr_add(seed, (uint64_t)j * (uint64_t)nnz + (uint64_t)t), k0, k1, rv);
        const uint64_t s = (uint64_t)rv[0] + ((uint64_t)rv[1] << 32);          // promote_uint_pair (util.hh:516-518)
        const uint32_t p = (uint32_t)t + (uint32_t)(s % (uint64_t)(d - t));    // sample from {t, ..., d-1}
        const uint32_t a = get(p), b = get((uint32_t)t);
        if (nnz > 1) { lp[nlog] = p; lv[nlog] = b; ++nlog; lp[nlog] = (uint32_t)t; lv[nlog] = a; ++nlog; }
        rows[t] = a;
        neg[t] = (rv[2] & 1u) ? 1 : 0;                                         // (rv[2] % 2 == 0) ? +1 : -1
    }
    for (int a = 1; a < nnz; ++a) {
        uint32_t key = rows[a]; uint8_t v = neg[a];
        int c = a - 1;
        for (; c >= 0 && rows[c] > key; --c) { rows[c + 1] = rows[c]; neg[c + 1] = neg[c]; }
        rows[c + 1] = key; neg[c + 1] = v;
    }
}

// ---- COO export (fill_sparse_unpacked) for parity tests and host callers ---------------------------------------------------------
// vector i (absolute long-axis index long_off + i) keeps the entries whose short-axis index lies in [short_off, short_off + short_sub)
__global__ void __launch_bounds__(256) saso_count_kernel(Ctr128 seed, uint32_t k0, uint32_t k1, int64_t long_off, int64_t long_sub, int nnz,
                                                         int64_t dim_major, int64_t short_off, int64_t short_sub, int* __restrict__ cnt) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < long_sub; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t rows[kSasoMaxNnz]; uint8_t neg[kSasoMaxNnz];
        saso_column(seed, k0, k1, long_off + i, nnz, dim_major, rows, neg);
        int c = 0;
        for (int t = 0; t < nnz; ++t) c += ((int64_t)rows[t] >= short_off && (int64_t)rows[t] < short_off + short_sub);
        cnt[i] = c;
    }
}
// exclusive scan of cnt[0..n) -> pos[0..n], single CTA
__global__ void __launch_bounds__(1024) scan_int_kernel(const int* __restrict__ cnt, int64_t n, int64_t* __restrict__ pos) {
    __shared__ int64_t wsum[32];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + threadIdx.x;
        int64_t v = i < n ? cnt[i] : 0, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int64_t y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int64_t w = wsum[threadIdx.x], xs = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int64_t y = __shfl_up_sync(0xffffffffu, xs, o); if (threadIdx.x >= o) xs += y; }
            wsum[threadIdx.x] = xs - w;
        }
        __syncthreads();
        const int64_t excl = carry + wsum[threadIdx.x >> 5] + x - v;
        if (i < n) pos[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) pos[n] = carry;
}
template <typename T>
__global__ void __launch_bounds__(256) saso_coo_kernel(Ctr128 seed, uint32_t k0, uint32_t k1, int64_t long_off, int64_t long_sub, int nnz,
                                                       int64_t dim_major, int64_t short_off, int64_t short_sub, const int64_t* __restrict__ pos,
                                                       T* __restrict__ vals, int64_t* __restrict__ idx_major, int64_t* __restrict__ idx_minor) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < long_sub; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t rows[kSasoMaxNnz]; uint8_t neg[kSasoMaxNnz];
        saso_column(seed, k0, k1, long_off + i, nnz, dim_major, rows, neg);
        int64_t o = pos[i];
        for (int t = 0; t < nnz; ++t) {
            const int64_t r = (int64_t)rows[t] - short_off;
            if (r >= 0 && r < short_sub) { idx_major[o] = r; idx_minor[o] = i; vals[o] = neg[t] ? (T)-1 : (T)1; ++o; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Axis::Long sparse operators (LASO, sparse_skops.hh:669-684): long-axis vector i draws vec_nnz iid uniform indices in [0, dim_major) - one Philox
// counter each, index = (rv[0] + 2^32 rv[1]) mod dim_major, sign from rv[2] (util.hh:521-541) - duplicates are merged into
// sqrt(count) * (sign of the first occurrence) (laso_merge_long_axis_vector_coo_data, :483-511) and the survivors sorted by index (:641-653).
// Returns the number of distinct indices; idx ascending, cnt their multiplicities, neg the first-occurrence signs.
// ------------------------------------------------------------------------------------------------
__device__ int laso_vector(const Ctr128& seed, uint32_t k0, uint32_t k1, int64_t i, int nnz, int64_t dim_major, int64_t* idx, int* cnt, uint8_t* neg) {
    int nd = 0;
    for (int t = 0; t < nnz; ++t) {
        uint32_t rv[4];
        philox4x32_10(ctr_add(seed, (uint64_t)i * (uint64_t)nnz + (uint64_t)t), k0, k1, rv);
        const uint64_t s = (uint64_t)rv[0] + ((uint64_t)rv[1] << 32);
        const int64_t ell = (int64_t)(s % (uint64_t)dim_major);
        int f = 0;
        for (; f < nd && idx[f] != ell; ++f) {}
        if (f < nd) cnt[f] += 1;
        else { idx[nd] = ell; cnt[nd] = 1; neg[nd] = (rv[2] & 1u) ? 1 : 0; ++nd; }
    }
    for (int a = 1; a < nd; ++a) {
        const int64_t key = idx[a]; const int c0 = cnt[a]; const uint8_t v = neg[a];
        int c = a - 1;
        for (; c >= 0 && idx[c] > key; --c) { idx[c + 1] = idx[c]; cnt[c + 1] = cnt[c]; neg[c + 1] = neg[c]; }
        idx[c + 1] = key; cnt[c + 1] = c0; neg[c + 1] = v;
    }
    return nd;
}
template <typename T>
__device__ __forceinline__ T laso_value(int count, uint8_t neg) {          // std::sqrt(c) * loc2scale[ell] in the working type (:506)
    T r;
    if constexpr (sizeof(T) == 8) r = sqrt((double)count); else r = sqrtf((float)count);
    return neg ? -r : r;
}
// vector v (absolute index vec_off + v) keeps the entries whose long-axis index lies in [long_off, long_off + long_sub)
__global__ void __launch_bounds__(128) laso_count_kernel(Ctr128 seed, uint32_t k0, uint32_t k1, int64_t vec_off, int64_t vec_sub, int nnz, int64_t dim_major,
                                                         int64_t long_off, int64_t long_sub, int* __restrict__ out) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < vec_sub; v += (int64_t)gridDim.x * blockDim.x) {
        int64_t idx[kSasoMaxNnz]; int cnt[kSasoMaxNnz]; uint8_t neg[kSasoMaxNnz];
        const int nd = laso_vector(seed, k0, k1, vec_off + v, nnz, dim_major, idx, cnt, neg);
        int c = 0;
        for (int t = 0; t < nd; ++t) c += (idx[t] >= long_off && idx[t] < long_off + long_sub);
        out[v] = c;
    }
}
template <typename T>
__global__ void __launch_bounds__(128) laso_coo_kernel(Ctr128 seed, uint32_t k0, uint32_t k1, int64_t vec_off, int64_t vec_sub, int nnz, int64_t dim_major,
                                                       int64_t long_off, int64_t long_sub, const int64_t* __restrict__ pos, T* __restrict__ vals,
                                                       int64_t* __restrict__ idx_major, int64_t* __restrict__ idx_minor) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < vec_sub; v += (int64_t)gridDim.x * blockDim.x) {
        int64_t idx[kSasoMaxNnz]; int cnt[kSasoMaxNnz]; uint8_t neg[kSasoMaxNnz];
        const int nd = laso_vector(seed, k0, k1, vec_off + v, nnz, dim_major, idx, cnt, neg);
        int64_t o = pos[v];
        for (int t = 0; t < nd; ++t) {
            const int64_t r = idx[t] - long_off;
            if (r >= 0 && r < long_sub) { idx_major[o] = r; idx_minor[o] = v; vals[o] = laso_value<T>(cnt[t], neg[t]); ++o; }
        }
    }
}
// B(d x n) = alpha * S[ro : ro + d, co : co + m] A + beta * B for a WIDE long-axis operator (its rows are the long-axis vectors): one CTA per
// sketch row regenerates the row's <= vec_nnz entries and gathers those rows of A; d * vec_nnz * n elements are touched in all.
template <typename T>
__global__ void __launch_bounds__(256) laso_apply_kernel(Ctr128 seed, uint32_t k0, uint32_t k1, int64_t ro, int nnz, int64_t dim_major, int64_t co,
                                                         int64_t m, int64_t n, const T* __restrict__ A, int64_t lda, double alpha, double beta,
                                                         T* __restrict__ B, int64_t ldb) {
    __shared__ int64_t s_idx[kSasoMaxNnz];
    __shared__ double s_w[kSasoMaxNnz];
    __shared__ int s_n;
    const int64_t r = blockIdx.x;
    if (threadIdx.x == 0) {
        int64_t idx[kSasoMaxNnz]; int cnt[kSasoMaxNnz]; uint8_t neg[kSasoMaxNnz];
        const int nd = laso_vector(seed, k0, k1, ro + r, nnz, dim_major, idx, cnt, neg);
        int k = 0;
        for (int t = 0; t < nd; ++t) {
            const int64_t j = idx[t] - co;
            if (j >= 0 && j < m) { s_idx[k] = j; s_w[k] = (double)laso_value<T>(cnt[t], neg[t]); ++k; }
        }
        s_n = k;
    }
    __syncthreads();
    const int k = s_n;
    for (int64_t c = threadIdx.x; c < n; c += blockDim.x) {
        double acc = 0.0;
        for (int t = 0; t < k; ++t) acc = fma(s_w[t], (double)A[s_idx[t] + c * lda], acc);
        double v = alpha * acc;
        if (beta != 0.0) v += beta * (double)B[r + c * ldb];
        B[r + c * ldb] = (T)v;
    }
}

static int saso_check(Ctx* ctx, int64_t n_rows, int64_t n_cols, int64_t vec_nnz, int major_axis) {
    RLB_REQUIRE(ctx, n_rows > 0);          // sparse_skops.hh:222-225
    RLB_REQUIRE(ctx, n_cols > 0);
    RLB_REQUIRE(ctx, vec_nnz > 0);
    RLB_REQUIRE(ctx, major_axis == RLB200_AXIS_SHORT || major_axis == RLB200_AXIS_LONG);
    // vec_nnz <= dim_major (sparse_skops.hh:241)
    RLB_REQUIRE(ctx, vec_nnz <= (major_axis == RLB200_AXIS_SHORT ? std::min(n_rows, n_cols) : std::max(n_rows, n_cols)));
    if (vec_nnz > kSasoMaxNnz) { ctx->err = "vec_nnz > 64 is not offered on the device"; return RLB200_ERR_UNSUPPORTED; }
    if (major_axis == RLB200_AXIS_SHORT) RLB_REQUIRE(ctx, std::min(n_rows, n_cols) < (1ll << 31));
    return 0;
}

static void saso_next_state(int64_t n_rows, int64_t n_cols, int64_t vec_nnz, uint32_t state[6]) {   // compute_next_state, sparse_skops.hh:302-312
    Ctr128 c;
    for (int i = 0; i < 4; ++i) c.v[i] = state[i];
    c = ctr_add(c, (uint64_t)(std::max(n_rows, n_cols) * vec_nnz));
    for (int i = 0; i < 4; ++i) state[i] = c.v[i];
}

template <typename T>
int fill_sparse_unpacked(Ctx* ctx, int64_t n_rows, int64_t n_cols, int64_t vec_nnz, int major_axis, int64_t sub_rows, int64_t sub_cols,
                         int64_t ro, int64_t co, int64_t* nnz_out, T* vals, int64_t* rows, int64_t* cols, uint32_t state[6]) {
    RLB_CHECK(saso_check(ctx, n_rows, n_cols, vec_nnz, major_axis));
    RLB_REQUIRE(ctx, sub_rows >= 0 && sub_cols >= 0 && ro >= 0 && co >= 0);
    RLB_REQUIRE(ctx, n_rows >= sub_rows + ro);      // sparse_skops.hh:575-576
    RLB_REQUIRE(ctx, n_cols >= sub_cols + co);
    RLB_REQUIRE(ctx, nnz_out != nullptr);
    if (major_axis == RLB200_AXIS_LONG) {
        // sparse_skops.hh:585-610, 669-704: the vectors run along the LONG axis and are indexed by the short one
        const bool short_rows = n_rows <= n_cols;
        const int64_t dim_major_l = std::max(n_rows, n_cols);
        const int64_t vec_off = short_rows ? ro : co, vec_sub = short_rows ? sub_rows : sub_cols;
        const int64_t lo = short_rows ? co : ro, ls = short_rows ? sub_cols : sub_rows;
        if (!vals || !rows || !cols) { *nnz_out = vec_nnz * vec_sub; return 0; }       // size query (:603-606): an upper bound for LASO
        Ctr128 seed_l;
        for (int i = 0; i < 4; ++i) seed_l.v[i] = state[i];
        *nnz_out = 0;
        if (vec_sub > 0) {
            ArenaScope as(ctx);
            int* cnt = as.take<int>(vec_sub); if (!cnt) return RLB200_ERR_ALLOC;
            int64_t* pos = as.take<int64_t>(vec_sub + 1); if (!pos) return RLB200_ERR_ALLOC;
            const int nb = (int)std::min<int64_t>((vec_sub + 127) / 128, (int64_t)ctx->num_sms * 16);
            LaunchScope ls_(ctx, RLB200_TIMER_FILL, 3);
            laso_count_kernel<<<nb, 128, 0, ctx->stream>>>(seed_l, state[4], state[5], vec_off, vec_sub, (int)vec_nnz, dim_major_l, lo, ls, cnt);
            scan_int_kernel<<<1, 1024, 0, ctx->stream>>>(cnt, vec_sub, pos);
            // major (long-axis) index = column of a wide operator, row of a tall one
            laso_coo_kernel<T><<<nb, 128, 0, ctx->stream>>>(seed_l, state[4], state[5], vec_off, vec_sub, (int)vec_nnz, dim_major_l, lo, ls, pos, vals,
                                                            short_rows ? cols : rows, short_rows ? rows : cols);
            RLB_CUDA_OK(ctx, cudaGetLastError());
            RLB_CUDA_OK(ctx, cudaMemcpyAsync(ctx->hbox, pos + vec_sub, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
            RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
            *nnz_out = *static_cast<int64_t*>(ctx->hbox);
        }
        Ctr128 nx = ctr_add(seed_l, (uint64_t)((vec_off + vec_sub) * vec_nnz));        // end_state after the last sampled vector
        for (int i = 0; i < 4; ++i) state[i] = nx.v[i];
        return 0;
    }
    const bool short_is_rows = n_rows <= n_cols;
    const int64_t dim_major = std::min(n_rows, n_cols);
    const int64_t short_off = short_is_rows ? ro : co, short_sub = short_is_rows ? sub_rows : sub_cols;
    const int64_t long_off = short_is_rows ? co : ro, long_sub = short_is_rows ? sub_cols : sub_rows;
    if (!vals || !rows || !cols) { *nnz_out = vec_nnz * long_sub; return 0; }   // size query (:603-606); state untouched
    Ctr128 seed;
    for (int i = 0; i < 4; ++i) seed.v[i] = state[i];
    *nnz_out = 0;
    if (long_sub > 0) {
        ArenaScope as(ctx);
        int* cnt = as.take<int>(long_sub); if (!cnt) return RLB200_ERR_ALLOC;
        int64_t* pos = as.take<int64_t>(long_sub + 1); if (!pos) return RLB200_ERR_ALLOC;
        const int nb = (int)std::min<int64_t>((long_sub + 255) / 256, (int64_t)ctx->num_sms * 16);
        LaunchScope ls(ctx, RLB200_TIMER_FILL, 3);
        saso_count_kernel<<<nb, 256, 0, ctx->stream>>>(seed, state[4], state[5], long_off, long_sub, (int)vec_nnz, dim_major, short_off, short_sub, cnt);
        scan_int_kernel<<<1, 1024, 0, ctx->stream>>>(cnt, long_sub, pos);
        saso_coo_kernel<T><<<nb, 256, 0, ctx->stream>>>(seed, state[4], state[5], long_off, long_sub, (int)vec_nnz, dim_major, short_off, short_sub, pos,
                                                        vals, short_is_rows ? rows : cols, short_is_rows ? cols : rows);
        RLB_CUDA_OK(ctx, cudaGetLastError());
        RLB_CUDA_OK(ctx, cudaMemcpyAsync(ctx->hbox, pos + long_sub, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        *nnz_out = *static_cast<int64_t*>(ctx->hbox);
    }
    // returned state = counter after the last sampled vector (sparse_skops.hh:613-614, 139)
    Ctr128 nx = ctr_add(seed, (uint64_t)((long_off + long_sub) * vec_nnz));
    for (int i = 0; i < 4; ++i) state[i] = nx.v[i];
    return 0;
}
template int fill_sparse_unpacked<double>(Ctx*, int64_t, int64_t, int64_t, int, int64_t, int64_t, int64_t, int64_t, int64_t*, double*, int64_t*, int64_t*, uint32_t*);
template int fill_sparse_unpacked<float>(Ctx*, int64_t, int64_t, int64_t, int, int64_t, int64_t, int64_t, int64_t, int64_t*, float*, int64_t*, int64_t*, uint32_t*);

// ---- plan: per-chunk CSR of the operator ----------------------------------------------------------------------------------------------
// chunk q covers columns [q*R, (q+1)*R) of the sub-operator (absolute column col0 + q*R + jl).
// key = (sketch row - ro) << 12 | jl << 1 | neg ; invalid = 0xFFFFFFFF.  ent[q*E + i] = key & 0xFFF, off[q*(d_pad+1) + r] = #keys < r<<12.
__global__ void __launch_bounds__(256) saso_plan_kernel(Ctr128 seed, uint32_t k0, uint32_t k1, int64_t col0, int64_t m_sub, int64_t d_full,
                                                        int64_t ro, int d_sub, int nnz, int R, int E2, int d_pad,
                                                        uint16_t* __restrict__ ent, uint16_t* __restrict__ off,
                                                        uint32_t* __restrict__ key_out = nullptr, int rows_per_group = 0, int ngroups = 0,
                                                        int sub = 0) {
    extern __shared__ uint32_t keys[];
    const int q = blockIdx.x;
    const int E = R * nnz;
    for (int i = E + threadIdx.x; i < E2; i += blockDim.x) keys[i] = 0xFFFFFFFFu;
    for (int jl = threadIdx.x; jl < R; jl += blockDim.x) {
        const int64_t j = (int64_t)q * R + jl;
        uint32_t rows[kSasoMaxNnz]; uint8_t neg[kSasoMaxNnz];
        if (j < m_sub) saso_column(seed, k0, k1, col0 + j, nnz, d_full, rows, neg);
        for (int t = 0; t < nnz; ++t) {
            uint32_t key = 0xFFFFFFFFu;
            if (j < m_sub) {
                const int64_t r = (int64_t)rows[t] - ro;
                if (r >= 0 && r < d_sub) {
                    uint32_t rr = (uint32_t)r;
                    if (sub > 0) {
                        // strip kernel: the `sub` lane groups of a warp share a range of sub * rows_per_group sketch rows and split it by
                        // row mod sub, so that their accumulator accesses fall into disjoint shared-memory banks; the key carries the
                        // row's index in that regrouped order (group = rr / rows_per_group)
                        const uint32_t span = (uint32_t)sub * (uint32_t)rows_per_group, w0 = rr / span, in = rr - w0 * span;
                        rr = w0 * span + (in % (uint32_t)sub) * (uint32_t)rows_per_group + in / (uint32_t)sub;
                    }
                    key = (rr << 12) | ((uint32_t)jl << 1) | neg[t];
                }
            }
            keys[jl * nnz + t] = key;
        }
    }
    __syncthreads();
    // bitonic sort of E2 (power of two) keys
    for (int k = 2; k <= E2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < E2; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const uint32_t a = keys[i], b = keys[l];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[l] = a; }
                }
            }
            __syncthreads();
        }
    }
    if (key_out) {
        // strip kernel format: the sorted 32-bit keys themselves (sketch row << 12 | source row << 1 | sign) and one offset per group of
        // rows_per_group sketch rows (ngroups + 1 values, padded to a multiple of 8)
        for (int i = threadIdx.x; i < E; i += blockDim.x) key_out[(int64_t)q * E + i] = keys[i];
        for (int gq = threadIdx.x; gq <= ngroups; gq += blockDim.x) {
            const int64_t r = (int64_t)gq * rows_per_group;      // (regrouped row indices run up to ngroups * rows_per_group)
            const uint32_t bound = (gq >= ngroups) ? 0xFFFFFFFFu : ((uint32_t)r << 12);
            int lo = 0, hi = E2;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] < bound) lo = mid + 1; else hi = mid; }
            off[(int64_t)q * (ngroups + 8) + gq] = (uint16_t)lo;
        }
        return;
    }
    for (int i = threadIdx.x; i < E; i += blockDim.x) ent[(int64_t)q * E + i] = (uint16_t)(keys[i] & 0xFFFu);
    for (int r = threadIdx.x; r <= d_pad; r += blockDim.x) {
        const uint32_t bound = (r >= d_sub) ? 0xFFFFFFFFu : ((uint32_t)r << 12);
        int lo = 0, hi = E2;   // first index with key >= bound
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] < bound) lo = mid + 1; else hi = mid; }
        off[(int64_t)q * (d_pad + 8) + r] = (uint16_t)lo;     // pitch d_pad + 8: every chunk's list starts 16-byte aligned
    }
}

// ---- apply ------------------------------------------------------------------------------------------------------------------------------
template <typename T, int RPT, int CW>
__global__ void __launch_bounds__(kSasoThreads, 1)
saso_apply_kernel(const T* __restrict__ A, int64_t lda, int64_t m, int n, int d_sub, int R, int E, int nchunks, int chunks_per_split,
                  const uint16_t* __restrict__ ent, const uint16_t* __restrict__ off, T* __restrict__ partial, int a_al16) {
    constexpr int EV = 16 / sizeof(T);
    constexpr int d_pad = kSasoThreads * RPT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);
    const int RP = R + EV;                       // row pitch of one staged column (keeps 16-byte alignment, shifts banks)
    // the chunk's CSR (entries + row offsets) is staged next to the tile: the list walk below reads shared memory only
    constexpr int OP = d_pad + 8;
    uint16_t* s_ent = reinterpret_cast<uint16_t*>(smem_raw + (size_t)2 * CW * RP * sizeof(T));
    uint16_t* s_off = s_ent + (size_t)2 * E;
    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * CW;
    const int split = blockIdx.y;
    const int q0 = split * chunks_per_split, q1 = min(nchunks, q0 + chunks_per_split);

    auto stage = [&](int buf, int q) {
        T* t = tile + (size_t)buf * CW * RP;
        const int64_t J0 = (int64_t)q * R;
        const int per_col = R / EV;
        for (int i = tid; i < CW * per_col; i += kSasoThreads) {
            const int c = i / per_col, jl = (i - c * per_col) * EV;
            int valid = (c0 + c < n) ? (int)max((int64_t)0, min((int64_t)EV, m - (J0 + jl))) : 0;
            const T* src = A + (J0 + jl) + (int64_t)(c0 + c) * lda;
            load_chunk<T>(t + c * RP + jl, valid > 0 ? src : A, valid, a_al16);
        }
        const uint16_t* ge = ent + (int64_t)q * E;
        for (int i = tid; i < E / 8; i += kSasoThreads) cp_async_16(s_ent + (size_t)buf * E + i * 8, ge + i * 8, true);
        const uint16_t* go = off + (int64_t)q * OP;
        for (int i = tid; i < OP / 8; i += kSasoThreads) cp_async_16(s_off + (size_t)buf * OP + i * 8, go + i * 8, true);
    };

    T acc[RPT][CW];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int c = 0; c < CW; ++c) acc[i][c] = (T)0;

    if (q0 < q1) stage(0, q0);
    cp_async_commit();
    for (int q = q0; q < q1; ++q) {
        const int buf = (q - q0) & 1;
        if (q + 1 < q1) stage(buf ^ 1, q + 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const uint16_t* op = s_off + (size_t)buf * OP + tid * RPT;
        int o[RPT + 1];
#pragma unroll
        for (int i = 0; i <= RPT; ++i) o[i] = op[i];
        const T* t = tile + (size_t)buf * CW * RP;
        const uint16_t* ep = s_ent + (size_t)buf * E;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            for (int e = o[i]; e < o[i + 1]; ++e) {
                const uint32_t w = ep[e];
                const int jl = w >> 1;
                const bool ng = w & 1u;
#pragma unroll
                for (int c = 0; c < CW; ++c) {
                    const T v = t[c * RP + jl];
                    acc[i][c] += ng ? -v : v;
                }
            }
        }
        __syncthreads();   // everyone is done with `buf` before it is refilled two iterations later
    }
    cp_async_wait<0>();
    T* P = partial + (int64_t)split * d_sub * n;
#pragma unroll
    for (int c = 0; c < CW; ++c) {
        if (c0 + c >= n) continue;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int r = tid * RPT + i;
            if (r < d_sub) P[r + (int64_t)(c0 + c) * d_sub] = acc[i][c];
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) saso_reduce_kernel(const T* __restrict__ partial, int splits, int d, int n, T alpha, T beta, T* __restrict__ B, int64_t ldb) {
    const int64_t total = (int64_t)d * n;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        T s = (T)0;
        for (int p = 0; p < splits; ++p) s += partial[(int64_t)p * total + e];
        T* b = B + (e % d) + (e / d) * ldb;
        T v = alpha * s;
        if (beta != (T)0) v += beta * (*b);
        *b = v;
    }
}

template <typename T, int RPT, int CW>
static int saso_apply_launch(Ctx* ctx, const T* A, int64_t lda, int64_t m, int n, int d_sub, int R, int E, int nchunks, const uint16_t* ent,
                             const uint16_t* off, T alpha, T beta, T* B, int64_t ldb) {
    constexpr int EV = 16 / sizeof(T);
    const int tiles = (n + CW - 1) / CW;
    int splits = (int)std::max<int64_t>(1, std::min<int64_t>(nchunks, (4ll * ctx->num_sms + tiles - 1) / tiles));
    const int cps = (nchunks + splits - 1) / splits;
    splits = (nchunks + cps - 1) / cps;
    ArenaScope as(ctx);
    T* partial = as.take<T>((size_t)splits * d_sub * n); if (!partial) return RLB200_ERR_ALLOC;
    const size_t smem = (size_t)2 * CW * (R + EV) * sizeof(T) + (size_t)2 * (E + kSasoThreads * RPT + 8) * sizeof(uint16_t);
    auto kern = saso_apply_kernel<T, RPT, CW>;
    RLB_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int al = (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (lda % EV == 0);
    LaunchScope ls(ctx, RLB200_TIMER_SKETCH, 2);
    kern<<<dim3(tiles, splits), kSasoThreads, smem, ctx->stream>>>(A, lda, m, n, d_sub, R, E, nchunks, cps, ent, off, partial, al);
    const int64_t total = (int64_t)d_sub * n;
    saso_reduce_kernel<T><<<(unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(partial, splits, d_sub, n, alpha, beta, B, ldb);
    RLB_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

// ---- apply, strip kernel -------------------------------------------------------------------------------------------------------------------
// A CTA owns CPS columns of A and ALL sketch rows, with its d x CPS accumulator tile in SHARED memory.  The 16 * (32 / CPS) lane
// groups each own a contiguous range of sketch rows; inside a group lane l owns column l % CPS.  The chunk's entries arrive sorted by
// sketch row (32-bit keys: row << 12 | source row << 1 | sign), so every group walks ONE flat list and adds tile[column][source row]
// into acc[row][column] - no per-row loops, no dynamic register indexing, and an accumulator is only ever touched by its owner lane in
// list order (deterministic, no atomics).  Chunks of R rows are fetched by one producer thread with cp.async.bulk (one copy per column
// + the chunk's keys and group offsets) into a double buffer guarded by full/empty mbarriers.  ncu on the generic kernel showed it
// issue-bound (~62 thread instructions per element); this form executes ~10x fewer.
__device__ __forceinline__ uint32_t ss_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ss_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void ss_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// +-v by flipping the sign bit with bit 0 of the key
__device__ __forceinline__ float saso_signed(float v, uint32_t k) { return __uint_as_float(__float_as_uint(v) ^ (k << 31)); }
__device__ __forceinline__ double saso_signed(double v, uint32_t k) {
    return __hiloint2double(__double2hiint(v) ^ (int)(k << 31), __double2loint(v));
}
constexpr int kStripThreads = 512;      // consumer threads (16 warps); one more warp produces
constexpr int kStripBufs = 2;           // chunk buffers (measured: 4 buffers of half the size are not faster - the loop is bound by the
                                        // shared-memory pipe, ~4 accesses per entry and lane group, not by load latency)

template <typename T, int CPS>
__global__ void __launch_bounds__(kStripThreads + 32, 1)
saso_strip_kernel(const T* __restrict__ A, int64_t lda, int n, int d_sub, int d_pad, int R, int E, int nchunks, int chunks_per_split,
                  const uint32_t* __restrict__ keys, const uint16_t* __restrict__ goff, T* __restrict__ partial) {
    constexpr int NG = 16 * (32 / CPS);            // lane groups = row ranges
    constexpr int OP = NG + 8;                     // group offsets per chunk (pitch of `goff`)
    constexpr int EV = 16 / sizeof(T);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NB = kStripBufs;
    __shared__ __align__(8) uint64_t bar_full[NB], bar_empty[NB];
    const int RP = R + EV;
    T* acc = reinterpret_cast<T*>(smem_raw);                                                     // [d_pad][CPS]
    T* tile = acc + (size_t)d_pad * CPS;                                                         // [NB][CPS][RP]
    uint32_t* s_key = reinterpret_cast<uint32_t*>(tile + (size_t)NB * CPS * RP);                 // [NB][E]
    uint16_t* s_off = reinterpret_cast<uint16_t*>(s_key + (size_t)NB * E);                       // [NB][OP]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c0 = blockIdx.x * CPS, split = blockIdx.y;
    const int q0 = split * chunks_per_split, q1 = min(nchunks, q0 + chunks_per_split);
    const int ncv = min(CPS, n - c0);                                                            // valid columns of this strip
    if (tid == 0) {
        for (int b = 0; b < NB; ++b) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ss_smem(&bar_full[b])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ss_smem(&bar_empty[b])), "n"(kStripThreads / 32));   // one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < d_pad * CPS; i += blockDim.x) acc[i] = (T)0;
    __syncthreads();

    if (warp == kStripThreads / 32) {
        // ---- producer
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)ncv * R * sizeof(T) + (uint32_t)E * 4u + (uint32_t)OP * 2u;
            for (int q = q0; q < q1; ++q) {
                const int i = q - q0, b = i % NB;
                if (i >= NB) ss_wait(ss_smem(&bar_empty[b]), (uint32_t)(((i / NB) - 1) & 1));
                const uint32_t bar = ss_smem(&bar_full[b]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                for (int c = 0; c < ncv; ++c)
                    ss_bulk(ss_smem(tile + ((size_t)b * CPS + c) * RP), A + (int64_t)q * R + (int64_t)(c0 + c) * lda, (uint32_t)(R * sizeof(T)), bar);
                ss_bulk(ss_smem(s_key + (size_t)b * E), keys + (int64_t)q * E, (uint32_t)E * 4u, bar);
                ss_bulk(ss_smem(s_off + (size_t)b * OP), goff + (int64_t)q * OP, (uint32_t)OP * 2u, bar);
            }
        }
    } else {
        // ---- consumers
        constexpr int SUB = 32 / CPS;
        const int col = lane % CPS, sub = lane / CPS, grp = warp * SUB + sub;
        // keys carry the row's index in the regrouped order (saso_plan_kernel): key row rr of group grp is sketch row
        // warp * SUB * rpg + sub + SUB * (rr - grp * rpg); the accumulator of (row, col) lives at acc[row * CPS + col]
        const int rpg = d_pad / NG;
        T* ac = acc + (warp * SUB * rpg + sub - SUB * grp * rpg) * CPS + col;
        for (int q = q0; q < q1; ++q) {
            const int i = q - q0, b = i % NB;
            ss_wait(ss_smem(&bar_full[b]), (uint32_t)((i / NB) & 1));
            const T* t = tile + ((size_t)b * CPS + col) * RP;
            const uint32_t* kp = s_key + (size_t)b * E;
            const int e1 = s_off[(size_t)b * OP + grp + 1];
            int e = s_off[(size_t)b * OP + grp];
            // four entries per step: their key / tile / accumulator loads are independent and in flight together; equal sketch rows are
            // adjacent in the sorted list, so a repeated accumulator index only needs the previous entry's updated value.  All indices are
            // 32-bit offsets into shared memory and the sign is applied by flipping the sign bit (ncu of the first version: issue-bound,
            // 65 % issue-active with the math pipe throttled by 64-bit address arithmetic).
            for (; e + 4 <= e1; e += 4) {
                uint32_t k[4]; T v[4]; int ai[4]; T x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) k[u] = kp[e + u];
#pragma unroll
                for (int u = 0; u < 4; ++u) { v[u] = t[(k[u] >> 1) & 0x7FFu]; ai[u] = (int)(k[u] >> 12) * (SUB * CPS); }
#pragma unroll
                for (int u = 0; u < 4; ++u) x[u] = ac[ai[u]];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (u > 0 && ai[u] == ai[u - 1]) x[u] = x[u - 1];
                    x[u] += saso_signed(v[u], k[u]);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) ac[ai[u]] = x[u];
            }
            for (; e < e1; ++e) {
                const uint32_t k = kp[e];
                const T v = t[(k >> 1) & 0x7FFu];
                const int ai = (int)(k >> 12) * (SUB * CPS);
                ac[ai] += saso_signed(v, k);
            }
            // release the buffer: one arrive per warp (no CTA-wide barrier, warps may run a few chunks apart)
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ss_smem(&bar_empty[b])) : "memory");
        }
    }
    __syncthreads();
    T* P = partial + (int64_t)split * d_sub * n;
    for (int i = tid; i < d_sub * ncv; i += blockDim.x) {
        const int c = i / d_sub, r = i - c * d_sub;
        P[r + (int64_t)(c0 + c) * d_sub] = acc[(size_t)r * CPS + c];
    }
}

// generic path (any alignment, any m): thread-owned sketch rows, saso_apply_kernel
template <typename T>
static int saso_apply_generic(Ctx* ctx, Ctr128 seed, const uint32_t* state, int64_t col0, int64_t m, int64_t S_rows, int64_t ro, int64_t d, int64_t vec_nnz,
                              int64_t n, const T* A, int64_t lda, T alpha, T beta, T* B, int64_t ldb) {
            int R = 2048;
            while (R > 256 && (int64_t)R * vec_nnz > 16384) R >>= 1;
            const int E = R * (int)vec_nnz;
            int E2 = 1; while (E2 < E) E2 <<= 1;
            const int nchunks = (int)((m + R - 1) / R);
            const int rpt = d <= 512 ? 1 : d <= 1024 ? 2 : d <= 4096 ? 8 : 32;
            const int d_pad = kSasoThreads * rpt;
            ArenaScope as(ctx);
            uint16_t* ent = as.take<uint16_t>((size_t)nchunks * E); if (!ent) return RLB200_ERR_ALLOC;
            uint16_t* off = as.take<uint16_t>((size_t)nchunks * (d_pad + 8)); if (!off) return RLB200_ERR_ALLOC;
            {
                RLB_CUDA_OK(ctx, cudaFuncSetAttribute(saso_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, E2 * 4));
                LaunchScope ls(ctx, RLB200_TIMER_SKETCH);
                saso_plan_kernel<<<nchunks, 256, E2 * 4, ctx->stream>>>(seed, state[4], state[5], col0, m, S_rows, ro, (int)d, (int)vec_nnz, R,
                                                                        E2, d_pad, ent, off);
                RLB_CUDA_OK(ctx, cudaGetLastError());
            }
            int rc;
            if (sizeof(T) == 4) {
                if (rpt == 1)      rc = saso_apply_launch<T, 1, 8>(ctx, A, lda, m, (int)n, (int)d, R, E, nchunks, ent, off, alpha, beta, B, ldb);
                else if (rpt == 2) rc = saso_apply_launch<T, 2, 8>(ctx, A, lda, m, (int)n, (int)d, R, E, nchunks, ent, off, alpha, beta, B, ldb);
                else if (rpt == 8) rc = saso_apply_launch<T, 8, 8>(ctx, A, lda, m, (int)n, (int)d, R, E, nchunks, ent, off, alpha, beta, B, ldb);
                else               rc = saso_apply_launch<T, 32, 2>(ctx, A, lda, m, (int)n, (int)d, R, E, nchunks, ent, off, alpha, beta, B, ldb);
            } else {
                if (rpt == 1)      rc = saso_apply_launch<T, 1, 4>(ctx, A, lda, m, (int)n, (int)d, R, E, nchunks, ent, off, alpha, beta, B, ldb);
                else if (rpt == 2) rc = saso_apply_launch<T, 2, 4>(ctx, A, lda, m, (int)n, (int)d, R, E, nchunks, ent, off, alpha, beta, B, ldb);
                else if (rpt == 8) rc = saso_apply_launch<T, 8, 4>(ctx, A, lda, m, (int)n, (int)d, R, E, nchunks, ent, off, alpha, beta, B, ldb);
                else               rc = saso_apply_launch<T, 32, 1>(ctx, A, lda, m, (int)n, (int)d, R, E, nchunks, ent, off, alpha, beta, B, ldb);
            }
            return rc;
}

// strip path: whole chunks of R rows of a 16-byte aligned A (saso_strip_kernel)
template <typename T, int CPS>
static int saso_apply_strips(Ctx* ctx, Ctr128 seed, const uint32_t* state, int64_t col0, int64_t m_full, int R, int64_t S_rows, int64_t ro, int64_t d,
                             int64_t vec_nnz, int64_t n, const T* A, int64_t lda, T alpha, T beta, T* B, int64_t ldb) {
    constexpr int NG = 16 * (32 / CPS);
    constexpr int EV = 16 / sizeof(T);
    const int nchunks = (int)(m_full / R);
    const int E = R * (int)vec_nnz;
    int E2 = 1; while (E2 < E) E2 <<= 1;
    const int rpg = (int)((d + NG - 1) / NG);          // sketch rows per lane group
    const int d_pad = rpg * NG, OP = NG + 8;
    ArenaScope as(ctx);
    uint32_t* keys = as.take<uint32_t>((size_t)nchunks * E); if (!keys) return RLB200_ERR_ALLOC;
    uint16_t* goff = as.take<uint16_t>((size_t)nchunks * OP); if (!goff) return RLB200_ERR_ALLOC;
    {
        RLB_CUDA_OK(ctx, cudaFuncSetAttribute(saso_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, E2 * 4));
        LaunchScope ls(ctx, RLB200_TIMER_SKETCH);
        saso_plan_kernel<<<nchunks, 256, E2 * 4, ctx->stream>>>(seed, state[4], state[5], col0, m_full, S_rows, ro, (int)d, (int)vec_nnz, R, E2, d_pad, nullptr,
                                                                goff, keys, rpg, NG, 32 / CPS);
        RLB_CUDA_OK(ctx, cudaGetLastError());
    }
    const int strips = (int)((n + CPS - 1) / CPS);
    int splits = (int)std::max<int64_t>(1, std::min<int64_t>(nchunks, (4ll * ctx->num_sms + strips - 1) / strips));
    const int cps = (nchunks + splits - 1) / splits;
    splits = (nchunks + cps - 1) / cps;
    T* partial = as.take<T>((size_t)splits * d * n); if (!partial) return RLB200_ERR_ALLOC;
    const size_t smem = (size_t)d_pad * CPS * sizeof(T) + (size_t)kStripBufs * (CPS * (R + EV) * sizeof(T) + E * 4 + OP * 2);
    auto kern = saso_strip_kernel<T, CPS>;
    RLB_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LaunchScope ls(ctx, RLB200_TIMER_SKETCH, 2);
    kern<<<dim3(strips, splits), kStripThreads + 32, smem, ctx->stream>>>(A, lda, (int)n, (int)d, d_pad, R, E, nchunks, cps, keys, goff, partial);
    const int64_t total = d * n;
    saso_reduce_kernel<T><<<(unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(partial, splits, (int)d, (int)n, alpha, beta, B, ldb);
    RLB_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

// B(d x n) = alpha * S[ro:ro+d, co:co+m] * A(m x n) + beta * B ; S ~ SparseDist(S_rows, S_cols, vec_nnz, Axis::Short) sampled at `state`.
// With a row shard set on the context, A holds rows [row_offset, row_offset + m) of the global matrix: the matching columns of S
// are used and B is sum-allreduced (beta is applied by the shard that owns row 0).  state <- S.next_state.
template <typename T>
int sketch_sparse_left(Ctx* ctx, int64_t S_rows, int64_t S_cols, int64_t vec_nnz, int64_t d, int64_t n, int64_t m, T alpha, int64_t ro,
                       int64_t co, const T* A, int64_t lda, T beta, T* B, int64_t ldb, uint32_t state[6]) {
    RLB_CHECK(saso_check(ctx, S_rows, S_cols, vec_nnz, RLB200_AXIS_SHORT));
    RLB_REQUIRE(ctx, d >= 0 && n >= 0 && m >= 0 && ro >= 0 && co >= 0);
    const bool sharded = ctx->m_global >= 0;
    const int64_t shard_off = sharded ? ctx->row_offset : 0;
    RLB_REQUIRE(ctx, S_rows >= d + ro);                                   // submatrix bounds (skge.hh / sparse_skops.hh:575-576)
    RLB_REQUIRE(ctx, S_cols >= (sharded ? ctx->m_global : m) + co);
    RLB_REQUIRE(ctx, lda >= m && ldb >= d);
    if (S_rows > S_cols) { ctx->err = "left sparse sketch with a tall operator is not offered on the device"; return RLB200_ERR_UNSUPPORTED; }
    if (d > kSasoThreads * 32) { ctx->err = "sketch dimension d > 16384 is not offered by the sparse sketch kernel"; return RLB200_ERR_UNSUPPORTED; }
    RLB_REQUIRE(ctx, n < (1ll << 31));
    Ctr128 seed;
    for (int i = 0; i < 4; ++i) seed.v[i] = state[i];
    if (d > 0 && n > 0) {
        if (m == 0) {
            // B <- beta * B
            const T bz = (sharded && shard_off != 0) ? (T)0 : beta;
            LaunchScope ls(ctx, RLB200_TIMER_SKETCH);
            saso_reduce_kernel<T><<<(unsigned)std::min<int64_t>((d * n + 255) / 256, 1184), 256, 0, ctx->stream>>>(B, 0, (int)d, (int)n, (T)0, bz, B, ldb);
            RLB_CUDA_OK(ctx, cudaGetLastError());
        } else {
            const T beta_eff = (sharded && shard_off != 0) ? (T)0 : beta;
            const int64_t col0 = co + shard_off;
            // whole chunks of a 16-byte aligned A go through the strip kernel (lanes = columns, cluster multicast of the strip);
            // the ragged tail (and unaligned / very wide-sketch cases) through the generic kernel.  The tail is applied first
            // (alpha, beta), the strips then accumulate with beta = 1: a fixed order.
            constexpr int EV = 16 / sizeof(T);
            // shared-memory budget of the strip kernel (<= 220 KB): accumulators d_pad x CPS, kStripBufs tile buffers CPS x (R + EV), keys E x 4 B each
            const int64_t acc_cap = (128 * 1024) / ((d + 127) / 128 * 128 * (int64_t)sizeof(T));       // columns per strip the accumulators allow
            const int cps_sel = acc_cap >= 8 ? 8 : acc_cap >= 4 ? 4 : acc_cap >= 2 ? 2 : 0;
            int Rs = 1024;
            while (Rs > 64 && ((int64_t)Rs * vec_nnz > 3072 || (int64_t)kStripBufs * cps_sel * (Rs + EV) * (int64_t)sizeof(T) > 68 * 1024)) Rs >>= 1;
            // measured (profiles/sec_sketch_sparse_*_r2b): the strip kernel wins for vec_nnz <= 4 (nnz 4: 68 vs 82 ms), the generic
            // one is kept for denser columns; RLB200_SASO_STRIPS=1 / RLB200_SASO_GENERIC=1 force either for experiments
            const bool strips_ok = cps_sel > 0 && (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (lda % EV == 0) && (int64_t)Rs * vec_nnz <= 3072 &&
                                   m >= 4 * (int64_t)Rs && getenv("RLB200_SASO_GENERIC") == nullptr && (vec_nnz <= 4 || getenv("RLB200_SASO_STRIPS") != nullptr);
            int rc = 0;
            if (strips_ok) {
                const int64_t m_full = (m / Rs) * Rs, tail = m - m_full;
                T beta_s = beta_eff;
                if (tail > 0) {
                    rc = saso_apply_generic<T>(ctx, seed, state, col0 + m_full, tail, S_rows, ro, d, vec_nnz, n, A + m_full, lda, alpha, beta_eff, B, ldb);
                    beta_s = (T)1;
                }
                if (rc >= 0) {
                    if (cps_sel == 8)      rc = saso_apply_strips<T, 8>(ctx, seed, state, col0, m_full, Rs, S_rows, ro, d, vec_nnz, n, A, lda, alpha, beta_s, B, ldb);
                    else if (cps_sel == 4) rc = saso_apply_strips<T, 4>(ctx, seed, state, col0, m_full, Rs, S_rows, ro, d, vec_nnz, n, A, lda, alpha, beta_s, B, ldb);
                    else                   rc = saso_apply_strips<T, 2>(ctx, seed, state, col0, m_full, Rs, S_rows, ro, d, vec_nnz, n, A, lda, alpha, beta_s, B, ldb);
                }
            } else {
                rc = saso_apply_generic<T>(ctx, seed, state, col0, m, S_rows, ro, d, vec_nnz, n, A, lda, alpha, beta_eff, B, ldb);
            }
            RLB_CHECK(rc);
        }
        if (sharded && ctx->allreduce) {
            if (ldb == d) {
                int rc = ctx->allreduce(ctx->allreduce_user, B, d * n, (int32_t)sizeof(T), ctx->stream);
                if (rc != 0) { ctx->err = "allreduce hook failed with code " + std::to_string(rc); return RLB200_ERR_COLLECTIVE; }
            } else {
                for (int64_t c = 0; c < n; ++c) {
                    int rc = ctx->allreduce(ctx->allreduce_user, B + c * ldb, d, (int32_t)sizeof(T), ctx->stream);
                    if (rc != 0) { ctx->err = "allreduce hook failed with code " + std::to_string(rc); return RLB200_ERR_COLLECTIVE; }
                }
            }
        }
    }
    saso_next_state(S_rows, S_cols, vec_nnz, state);
    return 0;
}
template int sketch_sparse_left<double>(Ctx*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, double, int64_t, int64_t, const double*, int64_t, double, double*, int64_t, uint32_t*);
template int sketch_sparse_left<float>(Ctx*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, float, int64_t, int64_t, const float*, int64_t, float, float*, int64_t, uint32_t*);

// ---- dense sketches -----------------------------------------------------------------------------------------------------------------------
// The operator is never materialised in HBM as a whole: column panels of S (left) / row panels (right) of at most `panel_bytes` are
// regenerated from the Philox state into a small ring buffer that stays L2-resident and are consumed at once by the tensor-pipe GEMM.
//
// left  (lskge3, skge.hh:155-203):  B(d x n) = alpha * S[ro:ro+d, co:co+m] * A(m x n) + beta * B
// right (rskge3, skge.hh:308-356):  B(m x d) = alpha * A(m x n) * S[ro:ro+n, co:co+d] + beta * B
// S ~ DenseDist(S_rows, S_cols, family, major_axis) sampled at `state`; state <- the full operator's next_state
// (DenseSkOp's constructor, dense_skops.hh:405-417).
static void dense_next_state_axis(int64_t n_rows, int64_t n_cols, int major_axis, uint32_t state[6]) {   // dense_skops.hh:169-182
    const int64_t mx = std::max(n_rows, n_cols), mn = std::min(n_rows, n_cols);
    const int64_t major = major_axis == RLB200_AXIS_LONG ? mx : mn, minor = major_axis == RLB200_AXIS_LONG ? mn : mx;
    Ctr128 c;
    for (int i = 0; i < 4; ++i) c.v[i] = state[i];
    c = ctr_add(c, (uint64_t)(((major + 3) / 4) * minor));
    for (int i = 0; i < 4; ++i) state[i] = c.v[i];
}

constexpr size_t kDensePanelBytes = (size_t)32 << 20;

// B(d x n, ldb) = alpha * Z(n x d, ldz)^T + beta * B
template <typename T>
__global__ void __launch_bounds__(256) sk_transpose_axpby_kernel(const T* __restrict__ Z, int64_t ldz, int64_t d, int64_t n, double alpha, double beta,
                                                                T* __restrict__ B, int64_t ldb) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t j0 = (int64_t)blockIdx.x * 32, i0 = (int64_t)blockIdx.y * 32;      // j: column of B (row of Z), i: row of B (column of Z)
    for (int r = ty; r < 32; r += 8) {
        const int64_t j = j0 + tx, i = i0 + r;
        tile[r][tx] = (j < n && i < d) ? (double)Z[j + i * ldz] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int64_t i = i0 + tx, j = j0 + r;
        if (i < d && j < n) {
            double v = alpha * tile[tx][r];
            if (beta != 0.0) v += beta * (double)B[i + j * ldb];
            B[i + j * ldb] = (T)v;
        }
    }
}

// the w x d column-major panel (ld = w) that holds rows j0 .. j0 + w of op(S)^T, op(S) = the d x m operator block applied from the left:
//   opS = NoTrans: the d x w block of S at (ro, co + j0), written ROW-major (= the natural, contiguous-write layout of a wide Long-axis operator);
//   opS = Trans  : op(S) = (the m x d block of S at (ro, co))^T, so the panel is the w x d block of S at (ro + j0, co), written COLUMN-major.
template <typename T>
static int fill_left_panel(Ctx* ctx, int64_t S_rows, int64_t S_cols, int family, int major_axis, int opS, int64_t d, int64_t w, int64_t ro, int64_t co,
                           int64_t j0, T* P, const uint32_t state[6]) {
    uint32_t st[6];
    std::memcpy(st, state, sizeof st);
    if (opS == 0) return fill_dense_unpacked<T>(ctx, S_rows, S_cols, family, major_axis, RLB200_LAYOUT_ROWMAJOR, d, w, ro, co + j0, P, st);
    return fill_dense_unpacked<T>(ctx, S_rows, S_cols, family, major_axis, RLB200_LAYOUT_COLMAJOR, w, d, ro + j0, co, P, st);
}

// sketch_general(ColMajor, NoTrans, NoTrans, d, n, m, alpha, S, ro, co, A, lda, beta, B, ldb) with a WIDE Axis::Long SparseSkOp
// (sparse_skops.hh:167-282, 669-684): state <- S.next_state = seed + min(S_rows, S_cols) * vec_nnz (:302-312).
template <typename T>
int sketch_sparse_left_laso(Ctx* ctx, int64_t S_rows, int64_t S_cols, int64_t vec_nnz, int64_t d, int64_t n, int64_t m, T alpha, int64_t ro,
                            int64_t co, const T* A, int64_t lda, T beta, T* B, int64_t ldb, uint32_t state[6]) {
    RLB_CHECK(saso_check(ctx, S_rows, S_cols, vec_nnz, RLB200_AXIS_LONG));
    RLB_REQUIRE(ctx, d >= 0 && n >= 0 && m >= 0 && ro >= 0 && co >= 0);
    RLB_REQUIRE(ctx, S_rows >= d + ro && S_cols >= m + co);
    RLB_REQUIRE(ctx, lda >= m && ldb >= d);
    if (ctx->m_global >= 0) { ctx->err = "long-axis sparse operators are not offered on a row-sharded context"; return RLB200_ERR_UNSUPPORTED; }
    if (S_rows > S_cols) { ctx->err = "left sparse sketch with a tall operator is not offered on the device"; return RLB200_ERR_UNSUPPORTED; }
    RLB_REQUIRE(ctx, d < (1ll << 31));
    Ctr128 seed;
    for (int i = 0; i < 4; ++i) seed.v[i] = state[i];
    if (d > 0 && n > 0) {
        LaunchScope ls(ctx, RLB200_TIMER_SKETCH);
        laso_apply_kernel<T><<<(unsigned)d, 256, 0, ctx->stream>>>(seed, state[4], state[5], ro, (int)vec_nnz, S_cols, co, m, n, A, lda, (double)alpha,
                                                                  (double)beta, B, ldb);
        RLB_CUDA_OK(ctx, cudaGetLastError());
    }
    Ctr128 nx = ctr_add(seed, (uint64_t)(std::min(S_rows, S_cols) * vec_nnz));
    for (int i = 0; i < 4; ++i) state[i] = nx.v[i];
    return 0;
}

template <typename T>
int sketch_dense_left(Ctx* ctx, int64_t S_rows, int64_t S_cols, int family, int major_axis, int64_t d, int64_t n, int64_t m, T alpha,
                      int64_t ro, int64_t co, const T* A, int64_t lda, T beta, T* B, int64_t ldb, uint32_t state[6], int opS) {
    RLB_REQUIRE(ctx, S_rows > 0 && S_cols > 0 && d >= 0 && n >= 0 && m >= 0 && ro >= 0 && co >= 0);
    RLB_REQUIRE(ctx, opS == 0 || opS == 1);
    const bool sharded = ctx->m_global >= 0;
    const int64_t shard_off = sharded ? ctx->row_offset : 0;
    // dims_before_op(d, m, opS) (skge.hh:118, 133-134)
    RLB_REQUIRE(ctx, (opS ? S_cols : S_rows) >= d + (opS ? co : ro));
    RLB_REQUIRE(ctx, (opS ? S_rows : S_cols) >= (sharded ? ctx->m_global : m) + (opS ? ro : co));
    RLB_REQUIRE(ctx, lda >= m && ldb >= d);
    if (d > 0 && n > 0) {
        const T beta0 = (sharded && shard_off != 0) ? (T)0 : beta;
        int64_t pc = std::max<int64_t>(64, (int64_t)(kDensePanelBytes / sizeof(T)) / std::max<int64_t>(d, 1));
        pc = std::min<int64_t>((pc / 64) * 64, std::max<int64_t>(m, 1));
        // tcgen05 int8 digit-slice engine for the long contraction: panels of four 16384-row accumulation chunks keep every SM busy
        const bool i8 = ctx->fp64_engine == RLB200_FP64_I8SLICES && m >= 16384 && d >= 64 && n >= 64;
        // (measured at d = 256, 2^22 x 2048 fp64: 65536-row panels 71.2 ms, 131072 67.5, 262144 63.2, 2^20 62.5 - the per-panel launches
        //  (fill, digit slicing, reduction) amortise; the two panel buffers are kept within 1 GiB)
        static const int64_t panel_rows_env = getenv("RLB200_DENSE_PANEL_ROWS") ? atoll(getenv("RLB200_DENSE_PANEL_ROWS")) : 0;
        int64_t panel_rows = panel_rows_env > 0 ? panel_rows_env : 262144;
        while (panel_rows_env <= 0 && panel_rows > 65536 && 2 * panel_rows * d * (int64_t)sizeof(T) > (1ll << 30)) panel_rows >>= 1;
        if (i8) pc = std::min<int64_t>(std::max<int64_t>(pc, panel_rows), ((m + 63) / 64) * 64);
        // Fused engine: the DATA matrix is the tall operand that is sliced inside the tensor-core kernel (read once, as fp64 / fp32), the
        // regenerated panels of S^T are the small operand; the kernel then produces (S A)^T, accumulated over the panels in scratch and
        // transposed into B at the end.
        const bool fused = i8 && ozaki2_tn_ok(ctx, std::min<int64_t>(pc, m), n, d, A, lda * (int64_t)sizeof(T));
        ArenaScope as(ctx);
        T* panel = as.take<T>((size_t)2 * d * pc); if (!panel) return RLB200_ERR_ALLOC;
        T* Zacc = nullptr;
        if (fused) { Zacc = as.take<T>((size_t)n * d); if (!Zacc) return RLB200_ERR_ALLOC; }
        if (m == 0) RLB_CHECK(gemm_nn<T>(ctx, d, n, 0, 0.0, panel, d, A, lda, (double)beta0, B, ldb));
        int buf = 0;
        for (int64_t j0 = 0; fused && j0 < m; j0 += pc, buf ^= 1) {
            const int64_t w = std::min(pc, m - j0);
            T* P = panel + (size_t)buf * d * pc;
            RLB_CHECK(fill_left_panel<T>(ctx, S_rows, S_cols, family, major_axis, opS, d, w, ro, co, shard_off + j0, P, state));
            RLB_CHECK(ozaki2_gemm_tn<T>(ctx, w, n, d, 1.0, A + j0, lda, P, w, j0 == 0 ? 0.0 : 1.0, Zacc, n));
        }
        if (fused) {
            LaunchScope ls(ctx, RLB200_TIMER_SKETCH);
            sk_transpose_axpby_kernel<T><<<dim3((unsigned)((n + 31) / 32), (unsigned)((d + 31) / 32)), 256, 0, ctx->stream>>>(Zacc, n, d, n, (double)alpha,
                                                                                                                          (double)beta0, B, ldb);
            RLB_CUDA_OK(ctx, cudaGetLastError());
            RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));      // Zacc is scratch of this scope
        }
        for (int64_t j0 = 0; !fused && j0 < m; j0 += pc, buf ^= 1) {
            const int64_t w = std::min(pc, m - j0);
            T* P = panel + (size_t)buf * d * pc;
            // the w x d column-major matrix op(S)_block^T with ld = w: the product is then the long-contraction
            // (split-K, whole-machine) form  B += (op(S)_block^T)^T A_block
            RLB_CHECK(fill_left_panel<T>(ctx, S_rows, S_cols, family, major_axis, opS, d, w, ro, co, shard_off + j0, P, state));
            if (i8) RLB_CHECK(ozaki_gemm_tn<T>(ctx, w, d, n, (double)alpha, P, w, A + j0, lda, j0 == 0 ? (double)beta0 : 1.0, B, ldb));
            else RLB_CHECK(gemm_tn<T>(ctx, w, d, n, (double)alpha, P, w, A + j0, lda, j0 == 0 ? (double)beta0 : 1.0, B, ldb, 0));
        }
        if (sharded && ctx->allreduce) {
            for (int64_t c = 0; c < (ldb == d ? 1 : n); ++c) {
                int rc = ctx->allreduce(ctx->allreduce_user, B + c * ldb, ldb == d ? d * n : d, (int32_t)sizeof(T), ctx->stream);
                if (rc != 0) { ctx->err = "allreduce hook failed with code " + std::to_string(rc); return RLB200_ERR_COLLECTIVE; }
            }
        }
    }
    dense_next_state_axis(S_rows, S_cols, major_axis, state);
    return 0;
}

template <typename T>
int sketch_dense_right(Ctx* ctx, int64_t S_rows, int64_t S_cols, int family, int major_axis, int64_t m, int64_t d, int64_t n, T alpha,
                       const T* A, int64_t lda, int64_t ro, int64_t co, T beta, T* B, int64_t ldb, uint32_t state[6], int opS) {
    RLB_REQUIRE(ctx, S_rows > 0 && S_cols > 0 && d >= 0 && n >= 0 && m >= 0 && ro >= 0 && co >= 0);
    RLB_REQUIRE(ctx, opS == 0 || opS == 1);
    // dims_before_op(n, d, opS) (skge.hh:271, 286-287)
    RLB_REQUIRE(ctx, (opS ? S_cols : S_rows) >= n + (opS ? co : ro));
    RLB_REQUIRE(ctx, (opS ? S_rows : S_cols) >= d + (opS ? ro : co));
    RLB_REQUIRE(ctx, lda >= m && ldb >= m);
    if (m > 0 && d > 0) {
        int64_t pr = std::max<int64_t>(64, (int64_t)(kDensePanelBytes / sizeof(T)) / std::max<int64_t>(d, 1));
        pr = std::min<int64_t>((pr / 64) * 64, std::max<int64_t>(n, 1));
        ArenaScope as(ctx);
        T* panel = as.take<T>((size_t)2 * d * pr); if (!panel) return RLB200_ERR_ALLOC;
        if (n == 0) RLB_CHECK(gemm_nn<T>(ctx, m, d, 0, 0.0, A, lda, panel, 1, (double)beta, B, ldb));
        int buf = 0;
        for (int64_t i0 = 0; i0 < n; i0 += pr, buf ^= 1) {
            const int64_t h = std::min(pr, n - i0);
            T* P = panel + (size_t)buf * d * pr;
            uint32_t st[6];
            std::memcpy(st, state, sizeof st);
            // rows i0 .. i0 + h of op(S) as an h x d column-major panel (ld = h): the h x d block of S at (ro + i0, co) written column-major,
            // or, for opS = Trans, the d x h block of S at (ro, co + i0) written row-major
            if (opS == 0) RLB_CHECK(fill_dense_unpacked<T>(ctx, S_rows, S_cols, family, major_axis, RLB200_LAYOUT_COLMAJOR, h, d, ro + i0, co, P, st));
            else RLB_CHECK(fill_dense_unpacked<T>(ctx, S_rows, S_cols, family, major_axis, RLB200_LAYOUT_ROWMAJOR, d, h, ro, co + i0, P, st));
            RLB_CHECK(gemm_nn<T>(ctx, m, d, h, (double)alpha, A + i0 * lda, lda, P, h, i0 == 0 ? (double)beta : 1.0, B, ldb));
        }
    }
    dense_next_state_axis(S_rows, S_cols, major_axis, state);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// sketch_general with every layout / transposition flag (RandBLAS/RandBLAS/skge.hh:859-905 left, :1031-1076 right; lskge3 :100-203,
// rskge3 :253-356).  The tall-product engines want the data matrix as a column-major (long x short) array and write a column-major
// result, so the other presentations are brought to that form:
//   * op(S) only changes which sub-block of the operator a panel regenerates and in which order it is written (fill_left_panel);
//   * a data matrix that arrives as the column-major (short x long) array - ColMajor + opA = Trans, or RowMajor + opA = NoTrans - is
//     transposed once into scratch (one extra read and write of A: HBM-bound, 3x the traffic of the direct form);
//   * a RowMajor result is the column-major transpose: computed in scratch, then written through the transposing axpby.
// ------------------------------------------------------------------------------------------------
// dst (cols x rows, ldd) = src (rows x cols, lds)^T, any size (the tile kernel's grid.y is bounded)
template <typename T>
static int transpose_big(Ctx* ctx, int64_t rows, int64_t cols, const T* src, int64_t lds, T* dst, int64_t ldd) {
    const int64_t step = (int64_t)1 << 20;
    for (int64_t c0 = 0; c0 < cols; c0 += step) {
        const int64_t w = std::min(step, cols - c0);
        RLB_CHECK(transpose<T>(ctx, rows, w, src + c0 * lds, lds, dst + c0, ldd));
    }
    return 0;
}

// op(A) (m x n) as a column-major m x n array: returns A itself or a transposed scratch copy
template <typename T>
static int data_as_colmajor(Ctx* ctx, ArenaScope& as, int layout, int opA, int64_t m, int64_t n, const T* A, int64_t lda, const T** out, int64_t* ld_out) {
    const bool colmajor = layout == RLB200_LAYOUT_COLMAJOR;
    // stored array, column-major view: ColMajor/NoTrans m x n; ColMajor/Trans n x m; RowMajor/NoTrans n x m; RowMajor/Trans m x n
    const bool direct = colmajor == (opA == 0);
    RLB_REQUIRE(ctx, lda >= std::max<int64_t>(1, direct ? m : n));          // skge.hh:136-143 / 289-296
    if (direct || m == 0 || n == 0) { *out = A; *ld_out = direct ? lda : std::max<int64_t>(m, 1); return 0; }
    T* At = as.take<T>((size_t)m * n); if (!At) return RLB200_ERR_ALLOC;
    RLB_CHECK(transpose_big<T>(ctx, n, m, A, lda, At, m));
    *out = At; *ld_out = m;
    return 0;
}

// run `inner(Bc, ldc, beta_c)` on a column-major r x c result: B itself (ColMajor), or scratch that is then written into the RowMajor B
template <typename T, typename F>
static int result_as_colmajor(Ctx* ctx, ArenaScope& as, int layout, int64_t r, int64_t c, T alpha, T beta, T* B, int64_t ldb, F&& inner) {
    if (layout == RLB200_LAYOUT_COLMAJOR) {
        RLB_REQUIRE(ctx, ldb >= std::max<int64_t>(1, r));
        return inner(B, ldb, alpha, beta);
    }
    RLB_REQUIRE(ctx, ldb >= std::max<int64_t>(1, c));                       // RowMajor r x c = column-major c x r with ld = ldb
    if (r == 0 || c == 0) return inner(B, std::max<int64_t>(r, 1), alpha, beta);
    T* Z = as.take<T>((size_t)r * c); if (!Z) return RLB200_ERR_ALLOC;
    RLB_CHECK(inner(Z, r, (T)1, (T)0));                                     // Z (r x c) = op(S) op(A)  or  op(A) op(S)
    LaunchScope ls(ctx, RLB200_TIMER_SKETCH);
    // B^T (c x r, ldb) = alpha * Z (r x c, ld r)^T + beta * B^T
    sk_transpose_axpby_kernel<T><<<dim3((unsigned)((r + 31) / 32), (unsigned)((c + 31) / 32)), 256, 0, ctx->stream>>>(Z, r, c, r, (double)alpha,
                                                                                                                  (double)beta, B, ldb);
    RLB_CUDA_OK(ctx, cudaGetLastError());
    RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));                   // Z is scratch of the caller's scope
    return 0;
}

template <typename T>
int sketch_general_dense_left(Ctx* ctx, int layout, int opS, int opA, int64_t d, int64_t n, int64_t m, T alpha, int64_t S_rows, int64_t S_cols,
                              int family, int major_axis, int64_t ro, int64_t co, const T* A, int64_t lda, T beta, T* B, int64_t ldb, uint32_t state[6]) {
    RLB_REQUIRE(ctx, layout == RLB200_LAYOUT_COLMAJOR || layout == RLB200_LAYOUT_ROWMAJOR);
    RLB_REQUIRE(ctx, (opS == 0 || opS == 1) && (opA == 0 || opA == 1));
    RLB_REQUIRE(ctx, d >= 0 && n >= 0 && m >= 0);
    if (layout == RLB200_LAYOUT_COLMAJOR && opA == 0) return sketch_dense_left<T>(ctx, S_rows, S_cols, family, major_axis, d, n, m, alpha, ro, co, A, lda, beta, B, ldb, state, opS);
    if (ctx->m_global >= 0) { ctx->err = "sketch_general on a row-sharded context: ColMajor with opA = NoTrans only"; return RLB200_ERR_UNSUPPORTED; }
    ArenaScope as(ctx);
    const T* Ac = nullptr; int64_t ldac = 0;
    RLB_CHECK(data_as_colmajor<T>(ctx, as, layout, opA, m, n, A, lda, &Ac, &ldac));
    return result_as_colmajor<T>(ctx, as, layout, d, n, alpha, beta, B, ldb, [&](T* Bc, int64_t ldc, T al, T be) {
        return sketch_dense_left<T>(ctx, S_rows, S_cols, family, major_axis, d, n, m, al, ro, co, Ac, ldac, be, Bc, ldc, state, opS);
    });
}

template <typename T>
int sketch_general_dense_right(Ctx* ctx, int layout, int opA, int opS, int64_t m, int64_t d, int64_t n, T alpha, const T* A, int64_t lda,
                               int64_t S_rows, int64_t S_cols, int family, int major_axis, int64_t ro, int64_t co, T beta, T* B, int64_t ldb, uint32_t state[6]) {
    RLB_REQUIRE(ctx, layout == RLB200_LAYOUT_COLMAJOR || layout == RLB200_LAYOUT_ROWMAJOR);
    RLB_REQUIRE(ctx, (opS == 0 || opS == 1) && (opA == 0 || opA == 1));
    RLB_REQUIRE(ctx, d >= 0 && n >= 0 && m >= 0);
    if (layout == RLB200_LAYOUT_COLMAJOR && opA == 0) return sketch_dense_right<T>(ctx, S_rows, S_cols, family, major_axis, m, d, n, alpha, A, lda, ro, co, beta, B, ldb, state, opS);
    ArenaScope as(ctx);
    const T* Ac = nullptr; int64_t ldac = 0;
    RLB_CHECK(data_as_colmajor<T>(ctx, as, layout, opA, m, n, A, lda, &Ac, &ldac));
    return result_as_colmajor<T>(ctx, as, layout, m, d, alpha, beta, B, ldb, [&](T* Bc, int64_t ldc, T al, T be) {
        return sketch_dense_right<T>(ctx, S_rows, S_cols, family, major_axis, m, d, n, al, Ac, ldac, ro, co, be, Bc, ldc, state, opS);
    });
}

// sketch_general with a SparseSkOp (short-axis / SASO) and every layout / transposition flag (skge.hh:907-960 left -> lskges :538-571;
// :1078-1131 right -> rskges :573-620).  A tall short-axis operator is the transpose of the wide one with swapped dimensions and the same
// seed (sparse_skops.hh:585-610: the index stream depends on (dim_major, dim_minor) only), so opS = Trans on a tall operator is the wide
// kernel with (S_rows, S_cols) and (ro, co) swapped.  The right sketch is the left sketch of the transposed problem, as in the reference:
// B = op(A) op(S)  <=>  B^T = op(S)^T op(A)^T, i.e. left(flipped layout, flipped opS, opA, d, m, n) on the same buffers.
template <typename T>
int sketch_general_sparse_left(Ctx* ctx, int layout, int opS, int opA, int64_t d, int64_t n, int64_t m, T alpha, int64_t S_rows, int64_t S_cols,
                               int64_t vec_nnz, int64_t ro, int64_t co, const T* A, int64_t lda, T beta, T* B, int64_t ldb, uint32_t state[6]) {
    RLB_REQUIRE(ctx, layout == RLB200_LAYOUT_COLMAJOR || layout == RLB200_LAYOUT_ROWMAJOR);
    RLB_REQUIRE(ctx, (opS == 0 || opS == 1) && (opA == 0 || opA == 1));
    RLB_REQUIRE(ctx, d >= 0 && n >= 0 && m >= 0);
    if (opS) {
        if (S_rows <= S_cols) { ctx->err = "sparse sketch_general: op(S) must be a wide short-axis operator (opS = Trans needs a tall S)"; return RLB200_ERR_UNSUPPORTED; }
        std::swap(S_rows, S_cols);
        std::swap(ro, co);
    }
    if (layout == RLB200_LAYOUT_COLMAJOR && opA == 0) return sketch_sparse_left<T>(ctx, S_rows, S_cols, vec_nnz, d, n, m, alpha, ro, co, A, lda, beta, B, ldb, state);
    if (ctx->m_global >= 0) { ctx->err = "sketch_general on a row-sharded context: ColMajor with opA = NoTrans only"; return RLB200_ERR_UNSUPPORTED; }
    ArenaScope as(ctx);
    const T* Ac = nullptr; int64_t ldac = 0;
    RLB_CHECK(data_as_colmajor<T>(ctx, as, layout, opA, m, n, A, lda, &Ac, &ldac));
    return result_as_colmajor<T>(ctx, as, layout, d, n, alpha, beta, B, ldb, [&](T* Bc, int64_t ldc, T al, T be) {
        return sketch_sparse_left<T>(ctx, S_rows, S_cols, vec_nnz, d, n, m, al, ro, co, Ac, ldac, be, Bc, ldc, state);
    });
}

template <typename T>
int sketch_general_sparse_right(Ctx* ctx, int layout, int opA, int opS, int64_t m, int64_t d, int64_t n, T alpha, const T* A, int64_t lda,
                                int64_t S_rows, int64_t S_cols, int64_t vec_nnz, int64_t ro, int64_t co, T beta, T* B, int64_t ldb, uint32_t state[6]) {
    RLB_REQUIRE(ctx, layout == RLB200_LAYOUT_COLMAJOR || layout == RLB200_LAYOUT_ROWMAJOR);
    RLB_REQUIRE(ctx, (opS == 0 || opS == 1) && (opA == 0 || opA == 1));
    if (ctx->m_global >= 0) { ctx->err = "right sparse sketch on a row-sharded context is not offered"; return RLB200_ERR_UNSUPPORTED; }
    const int flipped = layout == RLB200_LAYOUT_COLMAJOR ? RLB200_LAYOUT_ROWMAJOR : RLB200_LAYOUT_COLMAJOR;
    return sketch_general_sparse_left<T>(ctx, flipped, 1 - opS, opA, d, m, n, alpha, S_rows, S_cols, vec_nnz, ro, co, A, lda, beta, B, ldb, state);
}

#define INST(T)                                                                                                                                  \
    template int sketch_dense_left<T>(Ctx*, int64_t, int64_t, int, int, int64_t, int64_t, int64_t, T, int64_t, int64_t, const T*, int64_t, T, T*, int64_t, uint32_t*, int); \
    template int sketch_dense_right<T>(Ctx*, int64_t, int64_t, int, int, int64_t, int64_t, int64_t, T, const T*, int64_t, int64_t, int64_t, T, T*, int64_t, uint32_t*, int); \
    template int sketch_sparse_left_laso<T>(Ctx*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, T, int64_t, int64_t, const T*, int64_t, T, T*, int64_t, uint32_t*); \
    template int sketch_general_sparse_left<T>(Ctx*, int, int, int, int64_t, int64_t, int64_t, T, int64_t, int64_t, int64_t, int64_t, int64_t, const T*, int64_t, T, T*, int64_t, uint32_t*); \
    template int sketch_general_sparse_right<T>(Ctx*, int, int, int, int64_t, int64_t, int64_t, T, const T*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, T, T*, int64_t, uint32_t*); \
    template int sketch_general_dense_left<T>(Ctx*, int, int, int, int64_t, int64_t, int64_t, T, int64_t, int64_t, int, int, int64_t, int64_t, const T*, int64_t, T, T*, int64_t, uint32_t*); \
    template int sketch_general_dense_right<T>(Ctx*, int, int, int, int64_t, int64_t, int64_t, T, const T*, int64_t, int64_t, int64_t, int, int, int64_t, int64_t, T, T*, int64_t, uint32_t*);
INST(double)
INST(float)

}  // namespace rlb
