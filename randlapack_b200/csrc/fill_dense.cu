// fill_dense: materialise (a block of) a RandBLAS DenseDist sample on device.
// Counter -> entry layout follows dense::fill_dense_submat_impl
// (RandBLAS/RandBLAS/dense_skops.hh:96-167): the parent is imagined row-major with `n_cols`
// (= dim_major) columns, each row padded to a multiple of 4; parent entry (r, c) is lane c%4 of
// Philox(seed.ctr + r*ceil(n_cols/4) + c/4, seed.key).
//
// HBM-bound kernel: one thread per Philox counter (4 consecutive entries of a parent row), so a warp
// writes 128 consecutive entries (1 KB fp64 / 512 B fp32) of one parent row; algorithmic bytes =
// sizeof(T) per entry (write-only), no reads.
#include "common.cuh"
#include "philox.cuh"

namespace rlb {

template <typename T, int FAMILY>
__global__ void __launch_bounds__(256) fill_dense_kernel(T* __restrict__ out, int64_t nr, int64_t nc, int64_t r0, int64_t c0,
                                                         int64_t stride, int64_t blk0, int64_t nblk, int64_t rs, int64_t cs,
                                                         Ctr128 seed, uint32_t k0, uint32_t k1, T scale) {
    const int64_t total = nr * nblk;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / nblk, b = idx - row * nblk;
        const int64_t blk = blk0 + b;
        float rv[4];
        generate4<FAMILY>(ctr_add(seed, (uint64_t)((r0 + row) * stride + blk)), k0, k1, rv);
        const int64_t cbase = blk * 4 - c0;  // column (within the block) of lane 0
        T* orow = out + row * rs;
        if (cs == 1 && cbase >= 0 && cbase + 4 <= nc) {
            // contiguous fast path
            if (sizeof(T) == 8 && ((reinterpret_cast<uintptr_t>(orow + cbase) & 15) == 0)) {
                double2* p = reinterpret_cast<double2*>(orow + cbase);
                p[0] = make_double2((double)rv[0] * (double)scale, (double)rv[1] * (double)scale);
                p[1] = make_double2((double)rv[2] * (double)scale, (double)rv[3] * (double)scale);
            } else if (sizeof(T) == 4 && ((reinterpret_cast<uintptr_t>(orow + cbase) & 15) == 0)) {
                *reinterpret_cast<float4*>(orow + cbase) =
                    make_float4(rv[0] * (float)scale, rv[1] * (float)scale, rv[2] * (float)scale, rv[3] * (float)scale);
            } else {
#pragma unroll
                for (int l = 0; l < 4; ++l) orow[cbase + l] = (T)rv[l] * scale;
            }
        } else {
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                int64_t c = cbase + l;
                if (c >= 0 && c < nc) orow[c * cs] = (T)rv[l] * scale;
            }
        }
    }
}

// host launcher: `out` is written as out[row*rs + col*cs] over the nr x nc parent-order block.
template <typename T>
int fill_dense_launch(Ctx* ctx, int family, int64_t n_cols_parent, T* out, int64_t nr, int64_t nc, int64_t ptr, int64_t rs,
                      int64_t cs, const uint32_t state[6], uint32_t next_ctr[4]) {
    const int64_t stride = (n_cols_parent + 3) / 4;
    const int64_t r0 = ptr / n_cols_parent, c0 = ptr % n_cols_parent;
    const int64_t blk0 = c0 / 4, blk1 = (c0 + nc - 1) / 4;
    const int64_t nblk = blk1 - blk0 + 1;
    Ctr128 seed;
    for (int i = 0; i < 4; ++i) seed.v[i] = state[i];
    // Uniform family: entries scaled by sqrt(3) (dense_skops.hh:586, computed as (T)std::sqrt(3))
    T scale = family == RLB200_FAMILY_UNIFORM ? (T)1.7320508075688772 : (T)1;
    if (nr > 0 && nc > 0) {
        const int64_t total = nr * nblk;
        int64_t blocks = (total + 255) / 256;
        const int64_t cap = (int64_t)ctx->num_sms * 32;
        if (blocks > cap) blocks = cap;
        LaunchScope ls(ctx, RLB200_TIMER_FILL);
        if (family == RLB200_FAMILY_UNIFORM)
            fill_dense_kernel<T, RLB200_FAMILY_UNIFORM><<<(unsigned)blocks, 256, 0, ctx->stream>>>(
                out, nr, nc, r0, c0, stride, blk0, nblk, rs, cs, seed, state[4], state[5], scale);
        else
            fill_dense_kernel<T, RLB200_FAMILY_GAUSSIAN><<<(unsigned)blocks, 256, 0, ctx->stream>>>(
                out, nr, nc, r0, c0, stride, blk0, nblk, rs, cs, seed, state[4], state[5], scale);
        RLB_CUDA_OK(ctx, cudaGetLastError());
    }
    // next state: counter of the first block of the submatrix + nr strides (dense_skops.hh:164-166)
    Ctr128 nx = ctr_add(ctr_add(seed, (uint64_t)(r0 * stride + c0 / 4)), (uint64_t)(nr * stride));
    for (int i = 0; i < 4; ++i) next_ctr[i] = nx.v[i];
    return 0;
}

template int fill_dense_launch<double>(Ctx*, int, int64_t, double*, int64_t, int64_t, int64_t, int64_t, int64_t, const uint32_t*, uint32_t*);
template int fill_dense_launch<float>(Ctx*, int, int64_t, float*, int64_t, int64_t, int64_t, int64_t, int64_t, const uint32_t*, uint32_t*);

// DenseDist bookkeeping + fill_dense_unpacked semantics (dense_skops.hh:184-196, 560-603)
template <typename T>
int fill_dense_unpacked(Ctx* ctx, int64_t n_rows, int64_t n_cols, int family, int major_axis, int layout, int64_t sub_rows,
                        int64_t sub_cols, int64_t ro, int64_t co, T* buff, uint32_t state[6]) {
    RLB_REQUIRE(ctx, n_rows > 0 && n_cols > 0);
    RLB_REQUIRE(ctx, sub_rows >= 0 && sub_cols >= 0 && ro >= 0 && co >= 0);
    RLB_REQUIRE(ctx, n_rows >= sub_rows + ro);
    RLB_REQUIRE(ctx, n_cols >= sub_cols + co);
    RLB_REQUIRE(ctx, family == RLB200_FAMILY_GAUSSIAN || family == RLB200_FAMILY_UNIFORM);
    const int64_t mx = n_rows > n_cols ? n_rows : n_cols, mn = n_rows > n_cols ? n_cols : n_rows;
    const int64_t ma_len = (major_axis == RLB200_AXIS_LONG) ? mx : mn;
    const bool is_wide = n_rows < n_cols, fa_long = (major_axis == RLB200_AXIS_LONG);
    const bool nat_col = (is_wide && fa_long) ? false : (is_wide ? true : fa_long);
    const bool want_col = layout == RLB200_LAYOUT_NATURAL ? nat_col : (layout == RLB200_LAYOUT_COLMAJOR);
    int64_t nr, nc, ptr;
    if (nat_col) { nr = sub_cols; nc = sub_rows; ptr = ro + co * ma_len; }   // transpose viewed row-major
    else         { nr = sub_rows; nc = sub_cols; ptr = ro * ma_len + co; }
    // parent-order element (r, c) lands at: natural layout r*nc + c; flipped layout c*nr + r
    int64_t rs = (want_col == nat_col) ? nc : 1, cs = (want_col == nat_col) ? 1 : nr;
    uint32_t next[4];
    RLB_CHECK(fill_dense_launch<T>(ctx, family, ma_len, buff, nr, nc, ptr, rs, cs, state, next));
    for (int i = 0; i < 4; ++i) state[i] = next[i];
    return 0;
}
template int fill_dense_unpacked<double>(Ctx*, int64_t, int64_t, int, int, int, int64_t, int64_t, int64_t, int64_t, double*, uint32_t*);
template int fill_dense_unpacked<float>(Ctx*, int64_t, int64_t, int, int, int, int64_t, int64_t, int64_t, int64_t, float*, uint32_t*);

__global__ void philox_stream_kernel(uint32_t* __restrict__ out, int64_t n, Ctr128 seed, uint32_t k0, uint32_t k1) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t r[4];
        philox4x32_10(ctr_add(seed, (uint64_t)i), k0, k1, r);
        *reinterpret_cast<uint4*>(out + 4 * i) = make_uint4(r[0], r[1], r[2], r[3]);
    }
}

int philox_stream(Ctx* ctx, const uint32_t state[6], int64_t n, uint32_t* out) {
    RLB_REQUIRE(ctx, n >= 0);
    if (n == 0) return 0;
    Ctr128 seed;
    for (int i = 0; i < 4; ++i) seed.v[i] = state[i];
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)ctx->num_sms * 32) blocks = (int64_t)ctx->num_sms * 32;
    LaunchScope ls(ctx, RLB200_TIMER_FILL);
    philox_stream_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(out, n, seed, state[4], state[5]);
    RLB_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

}  // namespace rlb
