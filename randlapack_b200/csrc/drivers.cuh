// Host-side control flow of the sketch-and-factor stack on top of the sm_100a kernels:
// Stabilization (CholQRQ/...), RS, RF, QB, RSVD — same order of operations, same return codes and the
// same RNG-state advancement as the reference's `call`s, but device-resident and row-shardable.
#pragma once
#include "common.cuh"

namespace rlb {

// ---- kernel-level entry points (defined in the .cu files) -----------------------------------------
template <typename T>
int fill_dense_unpacked(Ctx* ctx, int64_t n_rows, int64_t n_cols, int family, int major_axis, int layout, int64_t sub_rows,
                        int64_t sub_cols, int64_t ro, int64_t co, T* buff, uint32_t state[6]);
int philox_stream(Ctx* ctx, const uint32_t state[6], int64_t n, uint32_t* out);
template <typename T>
int gemm_nn(Ctx* ctx, int64_t m, int64_t N, int64_t K, double alpha, const T* A, int64_t lda, const T* B, int64_t ldb, double beta,
            T* C, int64_t ldc);
template <typename T>
int gemm_nt(Ctx* ctx, int64_t m, int64_t N, int64_t K, double alpha, const T* A, int64_t lda, const T* B, int64_t ldb, double beta,
            T* C, int64_t ldc);
template <typename T>
int gemm_nn_inplace(Ctx* ctx, int64_t m, int64_t N, int64_t K, double alpha, T* X, int64_t ldx, const T* B, int64_t ldb,
                    bool b_upper_tri = false);
template <typename T>
int gemm_tn(Ctx* ctx, int64_t m, int64_t N1, int64_t N2, double alpha, const T* A, int64_t lda, const T* B, int64_t ldb, double beta,
            T* C, int64_t ldc, int upper_only, double* a_sumsq_out = nullptr);
template <typename T>
int sumsq(Ctx* ctx, const T* A, int64_t m, int64_t n, int64_t lda, double* partial_ws, double* out_dev);
int sumsq_ws_doubles(Ctx* ctx);
template <typename T>
int potrf_upper(Ctx* ctx, int k, T* A, int lda, int* info_dev);
template <typename T>
int trtri_upper(Ctx* ctx, int k, const T* R, int ldr, T* Rinv);
size_t svd_ws_bytes(int64_t n, int64_t k, size_t elem);
size_t plul_ws_bytes(Ctx* ctx, int64_t n);
template <typename T>
int plul(Ctx* ctx, int64_t m, int64_t n, T* A, int64_t lda, void* ws);
template <typename T>
int svd_tall(Ctx* ctx, int64_t n, int64_t k, T* B, int64_t ldb, T* S, T* W, void* ws, int* sweeps_out);

// ---- tall GEMMs on tcgen05 int8 tensor cores through exact digit slices (ozaki.cu) ------------------------
// C may alias A (in place: every row chunk is sliced before its rows are overwritten).  b_upper_tri: B is upper triangular
// (zeros below the diagonal are not multiplied).
template <typename T>
int ozaki_gemm_nn(Ctx* ctx, int64_t m, int64_t N, int64_t K, double alpha, const T* A, int64_t lda, const T* B, int64_t ldb, double beta,
                  T* C, int64_t ldc, bool b_upper_tri = false);
// x_sumsq_out (device scalar, optional): ||X||_F^2, accumulated in the exponent pass over X (no extra sweep).
// gram_out (optional, N2 x N2, ldg): Y^T Y computed in the same launches from the same digits of Y (tiles touching the upper triangle).
// upper_only: only the tiles touching the upper triangle of C are computed (Gram matrices; the rest of C is set to alpha * 0 + beta * C).
template <typename T>
int ozaki_gemm_tn(Ctx* ctx, int64_t m, int64_t N1, int64_t N2, double alpha, const T* X, int64_t ldx, const T* Y, int64_t ldy, double beta,
                  T* C, int64_t ldc, double* x_sumsq_out = nullptr, bool upper_only = false, T* gram_out = nullptr, int64_t ldg = 0);
void oz_cache_destroy(Ctx* ctx);
// ---- the same products with the digit slicing of the tall operand fused into the tensor-core kernel (ozaki_fused.cu) ----------
// C must not alias A.  *_ok: whether the fused engine takes the shape (otherwise call the staged engine above).
bool ozaki2_nn_ok(Ctx* ctx, int64_t m, int64_t N, int64_t K, const void* A, int64_t lda_bytes, const void* C);
bool ozaki2_tn_ok(Ctx* ctx, int64_t m, int64_t N1, int64_t N2, const void* X, int64_t ldx_bytes);
template <typename T>
int ozaki2_gemm_nn(Ctx* ctx, int64_t m, int64_t N, int64_t K, double alpha, const T* A, int64_t lda, const T* B, int64_t ldb, double beta,
                   T* C, int64_t ldc);
template <typename T>
int ozaki2_gemm_tn(Ctx* ctx, int64_t m, int64_t N1, int64_t N2, double alpha, const T* X, int64_t ldx, const T* Y, int64_t ldy, double beta,
                   T* C, int64_t ldc, double* x_sumsq_out = nullptr, T* gram_out = nullptr, int64_t ldg = 0, bool full_pairs = false);
void oz2_cache_destroy(Ctx* ctx);
// While alive, the first operand `A` of the tall products is known not to change: its row / column-chunk exponents are computed once
// and reused by every pass (outermost scope wins; nested scopes on the same or another pointer are no-ops).
struct OzConstScope {
    Ctx* ctx; bool owner;
    OzConstScope(Ctx* c, const void* A) : ctx(c), owner(c->oz_const_ptr == nullptr) {
        if (owner) { c->oz_const_ptr = A; invalidate(c); }
    }
    ~OzConstScope() { if (owner) { ctx->oz_const_ptr = nullptr; invalidate(ctx); } }
    static void invalidate(Ctx* c) { c->oz_row.valid = false; c->oz_col.valid = false; c->oz2_row.valid = false; c->oz2_col.valid = false; }
};

// ---- native collectives (comm.cu): NCCL communicator owned by the context, libnccl opened at run time -------------------
int comm_unique_id(unsigned char out[128], std::string* err);
int comm_init(Ctx* ctx, int nranks, int rank, const unsigned char id_bytes[128]);
void comm_destroy(Ctx* ctx);
bool comm_is_native(const Ctx* ctx);
bool comm_use_native(Ctx* ctx);

// ---- sketch-apply (sketch.cu) -----------------------------------------------------------------------
template <typename T>
int fill_sparse_unpacked(Ctx* ctx, int64_t n_rows, int64_t n_cols, int64_t vec_nnz, int major_axis, int64_t sub_rows, int64_t sub_cols,
                         int64_t ro, int64_t co, int64_t* nnz_out, T* vals, int64_t* rows, int64_t* cols, uint32_t state[6]);
template <typename T>
int sketch_sparse_left(Ctx* ctx, int64_t S_rows, int64_t S_cols, int64_t vec_nnz, int64_t d, int64_t n, int64_t m, T alpha, int64_t ro,
                       int64_t co, const T* A, int64_t lda, T beta, T* B, int64_t ldb, uint32_t state[6]);
template <typename T>
int sketch_dense_left(Ctx* ctx, int64_t S_rows, int64_t S_cols, int family, int major_axis, int64_t d, int64_t n, int64_t m, T alpha,
                      int64_t ro, int64_t co, const T* A, int64_t lda, T beta, T* B, int64_t ldb, uint32_t state[6], int opS = 0);
template <typename T>
int sketch_dense_right(Ctx* ctx, int64_t S_rows, int64_t S_cols, int family, int major_axis, int64_t m, int64_t d, int64_t n, T alpha,
                       const T* A, int64_t lda, int64_t ro, int64_t co, T beta, T* B, int64_t ldb, uint32_t state[6], int opS = 0);
// sketch_general with layout (RLB200_LAYOUT_COLMAJOR | _ROWMAJOR) and transposition flags (0 NoTrans, 1 Trans), skge.hh:859-905 / 1031-1076
template <typename T>
int sketch_general_dense_left(Ctx* ctx, int layout, int opS, int opA, int64_t d, int64_t n, int64_t m, T alpha, int64_t S_rows, int64_t S_cols,
                              int family, int major_axis, int64_t ro, int64_t co, const T* A, int64_t lda, T beta, T* B, int64_t ldb, uint32_t state[6]);
template <typename T>
int sketch_general_dense_right(Ctx* ctx, int layout, int opA, int opS, int64_t m, int64_t d, int64_t n, T alpha, const T* A, int64_t lda,
                               int64_t S_rows, int64_t S_cols, int family, int major_axis, int64_t ro, int64_t co, T beta, T* B, int64_t ldb, uint32_t state[6]);
template <typename T>
int sketch_sparse_left_laso(Ctx* ctx, int64_t S_rows, int64_t S_cols, int64_t vec_nnz, int64_t d, int64_t n, int64_t m, T alpha, int64_t ro,
                            int64_t co, const T* A, int64_t lda, T beta, T* B, int64_t ldb, uint32_t state[6]);
template <typename T>
int sketch_general_sparse_left(Ctx* ctx, int layout, int opS, int opA, int64_t d, int64_t n, int64_t m, T alpha, int64_t S_rows, int64_t S_cols,
                               int64_t vec_nnz, int64_t ro, int64_t co, const T* A, int64_t lda, T beta, T* B, int64_t ldb, uint32_t state[6]);
template <typename T>
int sketch_general_sparse_right(Ctx* ctx, int layout, int opA, int opS, int64_t m, int64_t d, int64_t n, T alpha, const T* A, int64_t lda,
                                int64_t S_rows, int64_t S_cols, int64_t vec_nnz, int64_t ro, int64_t co, T beta, T* B, int64_t ldb, uint32_t state[6]);
// B (cols x rows, ldb) = A (rows x cols, lda)^T (bqrrp.cu); cols / 32 < 65536
template <typename T>
int transpose(Ctx* ctx, int64_t rows, int64_t cols, const T* A, int64_t lda, T* B, int64_t ldb);

// ---- factorisation building blocks (factor.cu) -----------------------------------------------------
template <typename T>
int col_permute(Ctx* ctx, int64_t m, int64_t n, T* A, int64_t lda, const int64_t* perm_host);
template <typename T>
int tri_op(Ctx* ctx, int mode, int64_t rows, int64_t cols, const T* src, int64_t lds, T* dst, int64_t ldd);
template <typename T>
int potrf_blocked(Ctx* ctx, int64_t k, T* A, int64_t lda, int* info_host);
template <typename T>
int trsm_right_upper(Ctx* ctx, int64_t m, int64_t k, const T* R, int64_t ldr, T* X, int64_t ldx);
size_t qrcp_ws_bytes(int64_t n);
template <typename T>
int qr_small(Ctx* ctx, bool pivot, int64_t d, int64_t n, T* A, int64_t lda, int64_t* jpvt_dev, T* tau_dev, void* ws, int64_t stages = -1,
             double tol3z_in = 0.0);

template <typename T>
int make_unit_lower(Ctx* ctx, int64_t n, const T* src, int64_t lds, T* dst, int64_t ldd);
template <typename T>
int larft_from_gram(Ctx* ctx, int64_t k, const T* G, int64_t ldg, const T* tau, T* Tm, int64_t ldt);
template <typename T>
int set_upper_diag(Ctx* ctx, int64_t n, T* A, int64_t lda, T dval, bool add_to_diag);

// ---- arena: stack allocator for driver-level device buffers ---------------------------------------
void* arena_push(Ctx* ctx, size_t bytes);          // nullptr on failure (ctx->err set)
void arena_release(Ctx* ctx, size_t mark_total);   // pop back to a previous mark
size_t arena_mark(Ctx* ctx);
void arena_destroy(Ctx* ctx);

struct ArenaScope {
    Ctx* ctx; size_t mk;
    explicit ArenaScope(Ctx* c) : ctx(c), mk(arena_mark(c)) {}
    ~ArenaScope() { arena_release(ctx, mk); }
    template <typename T> T* take(size_t count) { return static_cast<T*>(arena_push(ctx, count * sizeof(T))); }
};

// ---- drivers ---------------------------------------------------------------------------------------
// rows_sharded: the m rows of A are this rank's block of a row-sharded matrix (Gram needs an allreduce);
// false for replicated operands (the n x k Omega).
template <typename T>
int stab_call(Ctx* ctx, int kind, int64_t m, int64_t k, T* A, bool cond_check, bool rows_sharded, int* chol_fail);
template <typename T>
int rs_call(Ctx* ctx, int64_t m, int64_t n, const T* A, int64_t k, T* Omega, T* work, uint32_t state[6], const rlb200_stack_opts& o);
template <typename T>
int rf_call(Ctx* ctx, int64_t m, int64_t n, const T* A, int64_t k, T* Q, uint32_t state[6], const rlb200_stack_opts& o);
template <typename T>
int qb_call(Ctx* ctx, int64_t m, int64_t n, T* A, int64_t* k, int64_t block_sz, T tol, T* Q, T* BT, T* Acpy, uint32_t state[6],
            const rlb200_stack_opts& o);
template <typename T>
int rsvd_call(Ctx* ctx, int64_t m, int64_t n, T* A, int64_t* k, T tol, T* U, T* S, T* V, T* Acpy, uint32_t state[6],
              const rlb200_stack_opts& o, int* qb_code);

// CQRRPT::call (rl_cqrrpt.hh:146-391), geqp3 flavour.  R_dev: n x n region (ldr >= n), J_dev: n pivots (1-based), *rank_out <- this->rank.
template <typename T>
int cqrrpt_call(Ctx* ctx, int64_t m, int64_t n, T* A, int64_t lda, T* R, int64_t ldr, int64_t* J_dev, T d_factor, T eps, int64_t nnz,
                int64_t* rank_out, uint32_t state[6]);

// CQRRT::call (rl_cqrrt.hh:91-297): unpivoted; returns 0, or 1 when the sketch's R has a zero diagonal entry or the Cholesky factorization fails.
template <typename T>
int cqrrt_call(Ctx* ctx, int64_t m, int64_t n, T* A, int64_t lda, T* R, int64_t ldr, T d_factor, int64_t nnz, int orthogonalization, int compute_Q,
               uint32_t state[6]);

// SYPS / SYRF / REVD2 (rl_syps.hh, rl_syrf.hh, rl_revd2.hh) on an explicit symmetric matrix of which the `uplo` triangle is read.
template <typename T>
int syps_call(Ctx* ctx, int uplo, int64_t m, const T* A, int64_t lda, int64_t k, int64_t passes, int64_t passes_per_stab, T* skop, T* work,
              uint32_t state[6]);
template <typename T>
int syrf_call(Ctx* ctx, int uplo, int64_t m, const T* A, int64_t lda, int64_t k, T* Q, T* work, uint32_t state[6], const rlb200_revd2_opts& o);
template <typename T>
int revd2_call(Ctx* ctx, int uplo, int64_t m, const T* A, int64_t lda, int64_t* k_io, int64_t k_cap, T tol, T* V, T* eigvals, uint32_t state[6],
               const rlb200_revd2_opts& o, T* err_out);

// BQRRP::call (rl_bqrrp.hh:154-665).  qrcp_wide: 0 luqr (default), 1 geqp3; qr_tall: 0 geqrf (default), 1 cholqr (+orhr_col), 2 geqrt.
// A_sk_ext != nullptr: BQRRP_GPU::call (rl_bqrrp_gpu.hh:152-942) - the d_ext x n sketch (leading dimension d_ext) is the caller's and is
// overwritten; d_factor and state are not used.
template <typename T>
int bqrrp_call(Ctx* ctx, int64_t m, int64_t n, T* A, int64_t lda, T d_factor, int64_t block_size, int qrcp_wide, int qr_tall, T* tau,
               int64_t* J_dev, int64_t* rank_out, uint32_t state[6], T* A_sk_ext = nullptr, int64_t d_ext = 0);

// hqrrp (rl_hqrrp.hh:811-1196): Householder QR with randomized pivoting, GEQP3 output format (A: R + reflectors, tau, J 1-based).
// qr_type (panel QR when panel_pivoting == 0): 0 / 1 Householder, 2 CholQR + Householder reconstruction.  Returns 0; 1 when the Cholesky
// factorization of a panel's Gram matrix fails (qr_type 2; the reference goes on with an unfactored panel there).
template <typename T>
int hqrrp_call(Ctx* ctx, int64_t m, int64_t n, T* A, int64_t lda, int64_t* J_dev, T* tau, int64_t nb_alg, int64_t pp, int panel_pivoting,
               int qr_type, uint32_t state[6]);

}  // namespace rlb
