/* oracle/librl_oracle.so — CPU restatement of the RandBLAS random-number / operator-entry
 * layer of the sketch-and-factor path, in plain C.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may load this library; the product (randlapack_b200/) never does.
 *
 * This is an independent restatement (not a copy): each function cites the reference lines whose
 * behaviour it reproduces.  Parity pins: Philox4x32-10 against the known-answer vectors the
 * reference ships (RandBLAS/test/basic_rng/r123_kat_vectors.txt:19-21); counter carries against
 * RandBLAS/test/basic_rng/test_r123.cc:735-796; operator entries and next-state arithmetic against
 * the real reference compiled in oracle/_ref (bit-exact, tests/test_oracle_*.py) and the golden
 * fixtures in tests/golden/ generated from it.
 *
 * Random123 (DEShawResearch/random123, HEAD at install time: install/install.sh:342) is not in the
 * reference tree; the Philox round function, u01/uneg11 and Box-Muller conventions below restate its
 * published algorithm (Salmon et al., SC'11; Random123 uniform.hpp / boxmuller.hpp docs).
 * The numerical drivers (RS/RF/QB/RSVD/CholQRQ/PLUL/HQRQ) are restated in oracle/rl_oracle.py on
 * top of the same LAPACK the reference build links.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "oracle_capi.h"

/* ---- Philox4x32-10 ------------------------------------------------------------------------- */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 128-bit little-endian counter += 64-bit step, wrapping (RNGState counter semantics,
 * RandBLAS/RandBLAS/base.hh:75-95; carry behaviour pinned by test_r123.cc:735-796). */
static void ctr_incr(uint32_t c[4], uint64_t step) {
    uint64_t carry = step;
    for (int i = 0; i < 4 && carry; ++i) {
        uint64_t s = (uint64_t)c[i] + (carry & 0xFFFFFFFFu);
        c[i] = (uint32_t)s;
        carry = (carry >> 32) + (s >> 32);
    }
}

/* Random123 conventions: u01 in (0,1], uneg11 in [-1,1], computed in float. */
static float u01f(uint32_t u)    { return (float)u * 0x1p-32f + 0x1p-33f; }
static float uneg11f(uint32_t u) { return (float)(int32_t)u * 0x1p-31f + 0x1p-32f; }

/* RandBLAS/RandBLAS/random_gen.hh:54-60 (host sincospif shim) + Random123 boxmuller(uint32,uint32). */
static void boxmuller_f(uint32_t u0, uint32_t u1, float* s_out, float* c_out) {
    const float PIf = 3.1415926535897932f;
    float x = uneg11f(u0);
    float s = sinf(PIf * x), c = cosf(PIf * x);
    float r = sqrtf(-2.f * logf(u01f(u1)));
    *s_out = s * r; *c_out = c * r;
}

/* r123ext::boxmul::generate / uneg11::generate (random_gen.hh:87-136,140-165): 4 floats per counter */
static void generate4(int family, const uint32_t ctr[4], const uint32_t key[2], float rv[4]) {
    uint32_t r[4];
    philox4x32_10(ctr, key, r);
    if (family == RL_FAMILY_GAUSSIAN) {
        boxmuller_f(r[0], r[1], &rv[0], &rv[1]);
        boxmuller_f(r[2], r[3], &rv[2], &rv[3]);
    } else {
        for (int i = 0; i < 4; ++i) rv[i] = uneg11f(r[i]);
    }
}

/* dense::fill_dense_submat_impl (RandBLAS/RandBLAS/dense_skops.hh:96-167): the parent is imagined
 * row-major with n_cols columns, every row padded to a multiple of 4; entry (r,c) of the parent is
 * lane c%4 of Philox(seed.ctr + r*ceil(n_cols/4) + c/4). Writes an n_srows x n_scols row-major
 * block (ld = n_scols) starting at flat offset ptr; returns ctr0 + n_srows*stride as next counter. */
#define DEFINE_FILL_SUBMAT(T, SUF)                                                                         \
static void fill_submat_##SUF(int family, int64_t n_cols, T* smat, int64_t n_srows, int64_t n_scols,        \
                              int64_t ptr, const uint32_t seed_ctr[4], const uint32_t key[2],               \
                              uint32_t next_ctr[4]) {                                                       \
    int64_t stride = (n_cols + 3) / 4;                                                                      \
    int64_t r0 = ptr / n_cols, c0 = ptr % n_cols;                                                           \
    _Pragma("omp parallel for schedule(static)")                                                            \
    for (int64_t row = 0; row < n_srows; ++row) {                                                           \
        int64_t blk = -1; float rv[4];                                                                      \
        for (int64_t j = 0; j < n_scols; ++j) {                                                             \
            int64_t c = c0 + j;                                                                             \
            if (c / 4 != blk) {                                                                             \
                blk = c / 4;                                                                                \
                uint32_t ctr[4]; memcpy(ctr, seed_ctr, 16);                                                 \
                ctr_incr(ctr, (uint64_t)((r0 + row) * stride + blk));                                       \
                generate4(family, ctr, key, rv);                                                            \
            }                                                                                               \
            smat[row * n_scols + j] = (T)rv[c % 4];                                                         \
        }                                                                                                   \
    }                                                                                                       \
    /* :164-166 — counter of the first block of the submatrix, plus n_srows strides */                      \
    memcpy(next_ctr, seed_ctr, 16);                                                                         \
    ctr_incr(next_ctr, (uint64_t)(r0 * stride + c0 / 4));                                                   \
    ctr_incr(next_ctr, (uint64_t)(n_srows * stride));                                                       \
}
DEFINE_FILL_SUBMAT(double, f64)
DEFINE_FILL_SUBMAT(float, f32)

/* DenseDist bookkeeping (dense_skops.hh:184-196, 319-331) */
static int natural_is_colmajor(int major_axis, int64_t n_rows, int64_t n_cols) {
    int is_wide = n_rows < n_cols, fa_long = (major_axis == RL_AXIS_LONG);
    if (is_wide && fa_long) return 0;
    if (is_wide) return 1;
    if (fa_long) return 1;
    return 0;
}

/* fill_dense_unpacked (dense_skops.hh:560-603): sub_rows x sub_cols block at (ro,co) of a sample of
 * DenseDist(n_rows,n_cols,family,major_axis), written in `layout`; Uniform scaled by sqrt(3) (:586). */
#define DEFINE_FILL_DENSE(T, SUF)                                                                           \
int rlo_fill_dense_##SUF(int64_t n_rows, int64_t n_cols, int family, int major_axis, int layout,            \
                         int64_t sub_rows, int64_t sub_cols, int64_t ro, int64_t co, T* buff,               \
                         uint32_t state[6]) {                                                               \
    if (n_rows <= 0 || n_cols <= 0) return RL_ERR_EXCEPTION;                                                \
    if (n_rows < sub_rows + ro || n_cols < sub_cols + co) return RL_ERR_EXCEPTION;                          \
    int64_t mx = n_rows > n_cols ? n_rows : n_cols, mn = n_rows > n_cols ? n_cols : n_rows;                 \
    int64_t ma_len = (major_axis == RL_AXIS_LONG) ? mx : mn;                                                \
    int nat_col = natural_is_colmajor(major_axis, n_rows, n_cols);                                          \
    int want_col = layout == RL_LAYOUT_NATURAL ? nat_col : (layout == RL_LAYOUT_COLMAJOR);                  \
    int64_t nr, nc, ptr;                                                                                    \
    if (nat_col) { nr = sub_cols; nc = sub_rows; ptr = ro + co * ma_len; }                                  \
    else         { nr = sub_rows; nc = sub_cols; ptr = ro * ma_len + co; }                                  \
    uint32_t next[4];                                                                                       \
    fill_submat_##SUF(family, ma_len, buff, nr, nc, ptr, state, state + 4, next);                           \
    if (family == RL_FAMILY_UNIFORM) {                                                                      \
        T s3 = (T)sqrt(3.0);                                                                                \
        for (int64_t i = 0; i < nr * nc; ++i) buff[i] *= s3;                                                \
    }                                                                                                       \
    if (want_col != nat_col) { /* out-of-place layout flip (:590-597) */                                    \
        T* w = (T*)malloc(sizeof(T) * (size_t)(sub_rows * sub_cols));                                       \
        memcpy(w, buff, sizeof(T) * (size_t)(sub_rows * sub_cols));                                         \
        for (int64_t i = 0; i < sub_rows; ++i)                                                              \
            for (int64_t j = 0; j < sub_cols; ++j) {                                                        \
                T v = nat_col ? w[i + j * sub_rows] : w[i * sub_cols + j];                                  \
                if (want_col) buff[i + j * sub_rows] = v; else buff[i * sub_cols + j] = v;                  \
            }                                                                                               \
        free(w);                                                                                            \
    }                                                                                                       \
    memcpy(state, next, 16);                                                                                \
    return 0;                                                                                               \
}
DEFINE_FILL_DENSE(double, f64)
DEFINE_FILL_DENSE(float, f32)

/* ---- flat helpers used by the tests ---------------------------------------------------------- */
const char* rlo_kind(void) { return "port"; }

int rlo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    philox4x32_10(ctr, key, out); return 0;
}
int rlo_ctr_incr(uint32_t ctr[4], uint64_t step) { ctr_incr(ctr, step); return 0; }
int rlo_boxmuller(uint32_t u0, uint32_t u1, float out[2]) { boxmuller_f(u0, u1, &out[0], &out[1]); return 0; }

/* raw Philox words for n consecutive counters (for the GPU integer-stream parity test) */
int rlo_philox_stream(const uint32_t ctr0[4], const uint32_t key[2], int64_t n, uint32_t* out) {
    for (int64_t i = 0; i < n; ++i) {
        uint32_t c[4]; memcpy(c, ctr0, 16); ctr_incr(c, (uint64_t)i);
        philox4x32_10(c, key, out + 4 * i);
    }
    return 0;
}

/* compute_next_state for a dense operator (dense_skops.hh:169-182) */
int rlo_dense_next_state(int64_t n_rows, int64_t n_cols, int major_axis, uint32_t state[6]) {
    int64_t mx = n_rows > n_cols ? n_rows : n_cols, mn = n_rows > n_cols ? n_cols : n_rows;
    int64_t major = (major_axis == RL_AXIS_LONG) ? mx : mn, minor = n_rows + n_cols - major;
    ctr_incr(state, (uint64_t)(((major + 3) / 4) * minor));
    return 0;
}

/* ---- SASO sampling (Axis::Short SparseDist) --------------------------------------------------------
 * Follows RandBLAS/RandBLAS/sparse_skops.hh:55-142 (_considerate_fisher_yates / repeated_fisher_yates: one Philox
 * call per draw, p = t + s % (dim_major - t), swap on an index array that is restored after every vector; vec_nnz == 1
 * delegates to sample_indices_iid_uniform, util.hh:520-542, which is the same arithmetic), :568-704
 * (fill_sparse_unpacked: per-vector insertion sort by the short-axis index, then the sub-block filter) and
 * :302-312 (next state).  Written with a real index array + undo pass, i.e. differently from the device kernel's
 * assignment log, so that the two implementations check each other.
 * major[i], minor[i], sign[i] for i < *nnz_out are relative to the sub-block; state <- counter after the last vector. */
int rlo_saso_coo(int64_t n_rows, int64_t n_cols, int64_t vec_nnz, int64_t sub_rows, int64_t sub_cols, int64_t ro, int64_t co,
                 int64_t* nnz_out, int64_t* rows, int64_t* cols, double* vals, uint32_t state[6]) {
    if (n_rows <= 0 || n_cols <= 0 || vec_nnz <= 0) return -1;
    const int short_is_rows = n_rows <= n_cols;
    const int64_t dim_major = short_is_rows ? n_rows : n_cols;
    if (vec_nnz > dim_major || n_rows < sub_rows + ro || n_cols < sub_cols + co) return -1;
    const int64_t short_off = short_is_rows ? ro : co, short_sub = short_is_rows ? sub_rows : sub_cols;
    const int64_t long_off = short_is_rows ? co : ro, long_sub = short_is_rows ? sub_cols : sub_rows;
    int64_t* idx = (int64_t*)malloc(sizeof(int64_t) * (size_t)dim_major);
    int64_t* piv = (int64_t*)malloc(sizeof(int64_t) * (size_t)vec_nnz);
    int64_t* smp = (int64_t*)malloc(sizeof(int64_t) * (size_t)vec_nnz);
    double* sg = (double*)malloc(sizeof(double) * (size_t)vec_nnz);
    for (int64_t i = 0; i < dim_major; ++i) idx[i] = i;
    uint32_t ctr[4]; memcpy(ctr, state, 16);
    ctr_incr(ctr, (uint64_t)(long_off * vec_nnz));
    int64_t out = 0;
    for (int64_t v = 0; v < long_sub; ++v) {
        for (int64_t t = 0; t < vec_nnz; ++t) {
            uint32_t rv[4];
            philox4x32_10(ctr, state + 4, rv);
            ctr_incr(ctr, 1);
            uint64_t s = (uint64_t)rv[0] + ((uint64_t)rv[1] << 32);
            int64_t p = t + (int64_t)(s % (uint64_t)(dim_major - t));
            piv[t] = p;
            int64_t tmp = idx[p]; idx[p] = idx[t]; idx[t] = tmp;
            smp[t] = idx[t];
            sg[t] = (rv[2] % 2 == 0) ? 1.0 : -1.0;
        }
        for (int64_t t = vec_nnz - 1; t >= 0; --t) { int64_t s = smp[t], p = piv[t]; idx[t] = idx[p]; idx[p] = s; }
        for (int64_t a = 1; a < vec_nnz; ++a) {            /* insertion sort by short-axis index */
            int64_t key = smp[a]; double kv = sg[a]; int64_t c = a - 1;
            for (; c >= 0 && smp[c] > key; --c) { smp[c + 1] = smp[c]; sg[c + 1] = sg[c]; }
            smp[c + 1] = key; sg[c + 1] = kv;
        }
        for (int64_t t = 0; t < vec_nnz; ++t) {
            int64_t mc = smp[t] - short_off;
            if (mc >= 0 && mc < short_sub) {
                if (short_is_rows) { rows[out] = mc; cols[out] = v; } else { cols[out] = mc; rows[out] = v; }
                vals[out] = sg[t];
                ++out;
            }
        }
    }
    free(idx); free(piv); free(smp); free(sg);
    *nnz_out = out;
    memcpy(state, ctr, 16);
    return 0;
}

/* SparseSkOp::next_state = compute_next_state (sparse_skops.hh:302-312), Axis::Short */
int rlo_saso_next_state(int64_t n_rows, int64_t n_cols, int64_t vec_nnz, uint32_t state[6]) {
    ctr_incr(state, (uint64_t)((n_rows > n_cols ? n_rows : n_cols) * vec_nnz));
    return 0;
}
