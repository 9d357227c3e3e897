/* rlb200.h — C-ABI of librlb200.so: a B200 (sm_100a) implementation of RandLAPACK's
 * sketch-and-factor hot path (RandBLAS dense sketch operators -> RS/RF rangefinder ->
 * CholQRQ -> QB -> RSVD).
 *
 * The reference (BallisticLA/RandLAPACK) has no FFI: its seam is C++ virtual dispatch on the
 * abstract bases Stabilization / RowSketcher / RangeFinder / QBalg / RSVDalg.  Each entry point
 * below therefore names the reference `call` (file:line, relative to the reference root) it
 * replaces; include/RandLAPACK_B200.hh wraps them in classes with the reference's constructor
 * arguments, public fields and `call` signatures (and, when RLB200_WITH_RANDLAPACK is defined,
 * deriving from the reference's own bases), see INTEGRATION.md.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types cross this boundary.
 *  - `_dev` entry points take DEVICE pointers (column-major, 64-bit dims) and enqueue on the
 *    context's stream; outputs live in caller-provided device buffers.
 *  - `_host` entry points take HOST pointers, stage through the device and block until done;
 *    they reproduce the reference's argument meaning for host callers.
 *  - return value: the reference's own int code for numeric events (see each function), or a
 *    negative RLB200_ERR_* for argument / CUDA / collective failures (never throws, never aborts).
 *  - RNG state: 6 x uint32 = Philox4x32 counter[4] (little-endian 128-bit) + key[2], in/out, with
 *    exactly the reference's advancement (RandBLAS/RandBLAS/base.hh:64-164).
 *  - there is NO CPU fallback: every compute entry point fails with RLB200_ERR_CUDA if no sm_100
 *    device is usable.
 */
#ifndef RLB200_H
#define RLB200_H
#include <stdint.h>
#include <stddef.h>

#if defined(__GNUC__)
#define RLB200_API __attribute__((visibility("default")))
#else
#define RLB200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define RLB200_ABI_VERSION 1

/* negative = infrastructure error; >= 0 = the reference's return code */
enum {
    RLB200_OK = 0,
    RLB200_ERR_ARG = -1,       /* what randlapack_require / randblas_require would have thrown on */
    RLB200_ERR_CUDA = -2,
    RLB200_ERR_ALLOC = -3,
    RLB200_ERR_COLLECTIVE = -4,
    RLB200_ERR_UNSUPPORTED = -5
};

/* Stabilization<T> implementations (RandLAPACK/comps/rl_orth.hh) */
enum { RLB200_STAB_PLUL = 0, RLB200_STAB_CHOLQRQ = 1, RLB200_STAB_HQRQ = 2 };
/* RandBLAS::ScalarDist / Axis / layout selectors (RandBLAS/RandBLAS/dense_skops.hh:209-347) */
enum { RLB200_FAMILY_GAUSSIAN = 0, RLB200_FAMILY_UNIFORM = 1 };
enum { RLB200_AXIS_LONG = 0, RLB200_AXIS_SHORT = 1 };
enum { RLB200_LAYOUT_NATURAL = 0, RLB200_LAYOUT_COLMAJOR = 1, RLB200_LAYOUT_ROWMAJOR = 2 };

typedef struct rlb200_ctx rlb200_ctx;

/* The canonical algorithm stack of test/drivers/test_rsvd.cc:68-93, flattened.
 * Value-compatible with oracle/oracle_capi.h:rl_stack_opts. */
typedef struct rlb200_stack_opts {
    int64_t passes_over_data;  /* RS::passes_over_data  (rl_rs.hh:45)  */
    int64_t passes_per_stab;   /* RS::passes_per_stab   (rl_rs.hh:46)  */
    int64_t block_sz;          /* RSVD::block_sz        (rl_rsvd.hh:44) */
    int32_t stab;              /* RS stabiliser,        RLB200_STAB_*  */
    int32_t orth_rf;           /* RF orthogonaliser                    */
    int32_t orth_qb;           /* QB re-orthogonaliser                 */
    int32_t cond_check;        /* bool: CholQRQ cond check (rl_orth.hh:88-93); RS/RF cond logging is not offered */
    int32_t orth_check;        /* bool: QB orthogonality_check (rl_qb.hh:199-207,236-244) */
    int32_t reserved;          /* bit 0: 1 = do NOT fold CholQR's triangular solve into the next product (default 0 = fold; same outputs) */
} rlb200_stack_opts;

/* The algorithm objects of test/drivers/test_revd2.cc:78-101, flattened (SYPS(p, q), SYRF(syps, orth), REVD2(syrf, error_est_p)). */
enum { RLB200_UPLO_UPPER = 0, RLB200_UPLO_LOWER = 1 };
typedef struct rlb200_revd2_opts {
    int64_t syps_passes;           /* SYPS::passes_over_data (rl_syps.hh:27) */
    int64_t syps_passes_per_stab;  /* SYPS::passes_per_stab  (rl_syps.hh:28) */
    int32_t orth;                  /* SYRF's orthogonaliser, RLB200_STAB_* (the reference's tests use HQRQ) */
    int32_t error_est_p;           /* REVD2::error_est_p     (rl_revd2.hh:81) */
} rlb200_revd2_opts;

/* Sum-allreduce hook for row-sharded operation (net-new; the reference is single-address-space).
 * Called on `count` elements of `elem_size` bytes at DEVICE pointer `buf`, stream-ordered on
 * `stream` (a cudaStream_t).  Return 0 on success.  NULL hook = single shard. */
typedef int (*rlb200_allreduce_fn)(void* user, void* buf, int64_t count, int32_t elem_size, void* stream);

/* ---- context ------------------------------------------------------------------------------- */
RLB200_API int rlb200_abi_version(void);
/* device: CUDA ordinal; stream: cudaStream_t to enqueue on (NULL = legacy default stream). */
RLB200_API int rlb200_create(rlb200_ctx** out, int device, void* stream);
RLB200_API int rlb200_destroy(rlb200_ctx* ctx);
RLB200_API const char* rlb200_last_error(const rlb200_ctx* ctx);
RLB200_API int rlb200_set_stream(rlb200_ctx* ctx, void* stream);
RLB200_API int rlb200_synchronize(rlb200_ctx* ctx);
/* Row-sharding: this process holds rows [row_offset, row_offset+m_local) of an m_global-row matrix. */
RLB200_API int rlb200_set_shard(rlb200_ctx* ctx, int64_t row_offset, int64_t m_global, rlb200_allreduce_fn fn, void* user);
/* Position of this shard among the row shards, needed by the TSQR orthogonaliser (HQRQ on a sharded iterate stacks the k x k
 * factors in rank order).  rlb200_comm_init sets it; hosts that use the hook declare it here. */
RLB200_API int rlb200_set_shard_rank(rlb200_ctx* ctx, int rank, int world);
/* Native data plane: an NCCL communicator owned by the context (libnccl is opened at run time; RLB200_NCCL_LIB overrides the name).
 * One rank calls rlb200_comm_unique_id and distributes the 128 bytes by any means (MPI, a file, torch.distributed); every rank then
 * calls rlb200_comm_init on its context (collective).  From then on the Gram / B^T / norm / R-factor sum-allreduces of the row-sharded
 * drivers are ncclAllReduce calls on the context's stream, and rlb200_set_shard(.., NULL, NULL) keeps using them. */
RLB200_API int rlb200_comm_unique_id(unsigned char id_out[128]);
RLB200_API int rlb200_comm_init(rlb200_ctx* ctx, int nranks, int rank, const unsigned char id[128]);
RLB200_API int rlb200_comm_destroy(rlb200_ctx* ctx);
/* kernels launched since creation / last reset (for bench.py's gpu_launches). */
RLB200_API int64_t rlb200_launch_count(rlb200_ctx* ctx, int reset);
/* CUDA-event timing of the kernels tagged `which` (see RLB200_TIMER_*), ms since last reset. */
enum { RLB200_TIMER_GEMM_NN = 0, RLB200_TIMER_GEMM_TN = 1, RLB200_TIMER_RIGHTMUL = 2, RLB200_TIMER_SMALL = 3, RLB200_TIMER_FILL = 4, RLB200_TIMER_SKETCH = 5,
       RLB200_TIMER_FACTOR = 6, RLB200_TIMER_I8_MMA_NN = 7, RLB200_TIMER_I8_MMA_TN = 8, RLB200_TIMER_I8_SLICE = 9, RLB200_TIMER_COUNT = 10 };
RLB200_API int rlb200_timers_enable(rlb200_ctx* ctx, int on);
RLB200_API int rlb200_timer_read(rlb200_ctx* ctx, int which, double* ms, int64_t* launches, int reset);

/* Per-phase wall-clock times of the QR drivers, the reference's public `times` vectors (microseconds, same order and length:
 * CQRRPT 8 entries rl_cqrrpt.hh:371-384, CQRRT 10 entries rl_cqrrt.hh:279-282, BQRRP 10 entries rl_bqrrp.hh:582-584).  Recorded only while
 * enabled (every phase boundary then synchronises the stream).  rlb200_get_phase_times returns the number of entries of the last call. */
RLB200_API int rlb200_set_phase_timing(rlb200_ctx* ctx, int on);
/* BQRRP's public `tol` field (rl_bqrrp.hh:141, used at :422 to cut the block rank where |R_sk(i,i)| / |R_sk(0,0)| < tol); 0 = the
 * constructor default (machine epsilon of the working type).  Applies to the following rlb200_bqrrp_* calls on this context.  The fields
 * `internal_nb` and `apply_trans_q` (:138, :149) select LAPACK blockings of the same factorization and have no counterpart here. */
RLB200_API int rlb200_set_bqrrp_tol(rlb200_ctx* ctx, double tol);
/* CQRRPT's public `qrcp` field (rl_cqrrpt.hh:41, used at :230-247): the QRCP of the sketch by lapack::geqp3 (default) or by BQRRP with the
 * reference's block ratio (n <= 2000: 1, n <= 8000: 1/2, else 1/32; BQRRP(false, n * ratio).call(d, n, A_hat, d, 1.0, tau, J, state) - the RNG
 * state advances through BQRRP's own sketch), or by hqrrp(d, n, A_hat, d, J, tau, nb_alg, oversampling, panel_pivoting, use_cholqr, state)
 * (:230-231; the state advances through hqrrp's uniform operator).  bqrrp and hqrrp: single shard.  Applies to the following rlb200_cqrrpt_* calls. */
enum { RLB200_CQRRPT_QRCP_GEQP3 = 0, RLB200_CQRRPT_QRCP_BQRRP = 1, RLB200_CQRRPT_QRCP_HQRRP = 2 };
RLB200_API int rlb200_set_cqrrpt_qrcp(rlb200_ctx* ctx, int qrcp);
/* CQRRPT's public HQRRP fields nb_alg, oversampling, panel_pivoting, use_cholqr (rl_cqrrpt.hh:134-137; constructor defaults 64, 10, 1, 0 at
 * :60-63), used when qrcp = hqrrp. */
RLB200_API int rlb200_set_cqrrpt_hqrrp_opts(rlb200_ctx* ctx, int64_t nb_alg, int64_t oversampling, int panel_pivoting, int use_cholqr);
/* CQRRPT's public `orthogonalization` field (rl_cqrrpt.hh:139-142, 343-368): R keeps the Cholesky factor (the preconditioning is not undone)
 * and, when rank < n, the trailing n - rank columns of A are completed to an orthonormal set: Gaussian columns (DenseDist(m, n - rank) drawn
 * at the current state, which - as in the reference - does not advance), projected against Q and orthogonalized by Householder QR.
 * n - rank <= 256, single shard. */
RLB200_API int rlb200_set_cqrrpt_orthogonalization(rlb200_ctx* ctx, int on);
RLB200_API int rlb200_get_phase_times(rlb200_ctx* ctx, long long* out_us, int cap);

/* ---- device memory helpers for host-pointer callers (the C++ adapters in RandLAPACK_B200.hh stage through these so that
 *      they need no CUDA headers).  Copies are stream-ordered on the context's stream; d2h blocks until complete. */
RLB200_API int rlb200_dev_alloc(rlb200_ctx* ctx, size_t bytes, void** out_dev);
RLB200_API int rlb200_dev_free(rlb200_ctx* ctx, void* dev);
RLB200_API int rlb200_copy_h2d(rlb200_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
RLB200_API int rlb200_copy_d2h(rlb200_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

/* ---- a1: Philox4x32-10 stream (r123::Philox4x32 via RandBLAS/RandBLAS/base.hh:53) -----------
 * out_dev[4*i .. 4*i+3] = Philox(counter = state.counter + i, key = state.key), i in [0,n). */
RLB200_API int rlb200_philox_stream_dev(rlb200_ctx* ctx, const uint32_t state[6], int64_t n, uint32_t* out_dev);

/* ---- a2-a4: RandBLAS::fill_dense / fill_dense_unpacked (dense_skops.hh:560-603, 620-655) ----
 * Writes the sub_rows x sub_cols block at (ro, co) of the sample of
 * DenseDist(n_rows, n_cols, family, major_axis) defined by `state`, in `layout`
 * (NATURAL = D.natural_layout), to buff_dev with leading dimension = the block's own extent.
 * state <- the reference's returned next state. */
RLB200_API int rlb200_fill_dense_f64_dev(rlb200_ctx* ctx, int64_t n_rows, int64_t n_cols, int family, int major_axis, int layout,
                              int64_t sub_rows, int64_t sub_cols, int64_t ro, int64_t co, double* buff_dev, uint32_t state[6]);
RLB200_API int rlb200_fill_dense_f32_dev(rlb200_ctx* ctx, int64_t n_rows, int64_t n_cols, int family, int major_axis, int layout,
                              int64_t sub_rows, int64_t sub_cols, int64_t ro, int64_t co, float* buff_dev, uint32_t state[6]);

/* ---- a6: RandBLAS::fill_sparse / fill_sparse_unpacked (RandBLAS/RandBLAS/sparse_skops.hh:568-704, 746) ---------------------
 * COO triplets (vals, rows, cols; 0-based, relative to the block) of the sub_rows x sub_cols block at (ro, co) of the sample of
 * SparseDist(n_rows, n_cols, vec_nnz, major_axis) defined by `state`, in the reference's order (one long-axis vector after the
 * other, short-axis indices ascending within a vector).  *nnz_out (HOST) receives the number of triplets; with NULL output
 * arrays only the size bound vec_nnz * (#long-axis vectors) is returned and `state` is untouched (the reference's size query).
 * state <- the reference's returned state (counter after the last sampled vector).  major_axis = RLB200_AXIS_SHORT (SASO, the default:
 * vec_nnz distinct short-axis indices per long-axis position, values +-1) or RLB200_AXIS_LONG (LASO, :669-704: each of the min(n_rows, n_cols)
 * long-axis vectors draws vec_nnz indices with replacement, duplicates merged into sqrt(count) * first sign, long-axis indices ascending
 * within a vector; the triplet count after merging is only known after the call). */
RLB200_API int rlb200_fill_sparse_f64_dev(rlb200_ctx* ctx, int64_t n_rows, int64_t n_cols, int64_t vec_nnz, int major_axis,
                               int64_t sub_rows, int64_t sub_cols, int64_t ro, int64_t co, int64_t* nnz_out, double* vals_dev,
                               int64_t* rows_dev, int64_t* cols_dev, uint32_t state[6]);
RLB200_API int rlb200_fill_sparse_f32_dev(rlb200_ctx* ctx, int64_t n_rows, int64_t n_cols, int64_t vec_nnz, int major_axis,
                               int64_t sub_rows, int64_t sub_cols, int64_t ro, int64_t co, int64_t* nnz_out, float* vals_dev,
                               int64_t* rows_dev, int64_t* cols_dev, uint32_t state[6]);

/* ---- a7: sketch_general(ColMajor, NoTrans, NoTrans, d, n, m, alpha, S, ro_s, co_s, A, lda, beta, B, ldb) with a SparseSkOp
 *      (RandBLAS/RandBLAS/skge.hh:538-571 lskges; call site RandLAPACK/drivers/rl_cqrrpt.hh:214-221).
 * S = SparseSkOp(SparseDist(S_rows, S_cols, vec_nnz, Axis::Short), state), wide (S_rows <= S_cols), regenerated on device from
 * the Philox state (never passed in).  B(d x n) = alpha * S[ro_s:ro_s+d, co_s:co_s+m] * A(m x n) + beta * B.
 * state <- S.next_state (sparse_skops.hh:302-312).  Row-sharded contexts use the shard's columns of S and allreduce B. */
RLB200_API int rlb200_sketch_sparse_left_f64_dev(rlb200_ctx* ctx, int64_t S_rows, int64_t S_cols, int64_t vec_nnz, int64_t d, int64_t n,
                                      int64_t m, double alpha, int64_t ro_s, int64_t co_s, const double* A_dev, int64_t lda, double beta,
                                      double* B_dev, int64_t ldb, uint32_t state[6]);
RLB200_API int rlb200_sketch_sparse_left_f32_dev(rlb200_ctx* ctx, int64_t S_rows, int64_t S_cols, int64_t vec_nnz, int64_t d, int64_t n,
                                      int64_t m, float alpha, int64_t ro_s, int64_t co_s, const float* A_dev, int64_t lda, float beta,
                                      float* B_dev, int64_t ldb, uint32_t state[6]);

/* The same call with a WIDE Axis::Long SparseSkOp (LASO, sparse_skops.hh:167-282, 669-684): every ROW of S holds vec_nnz iid uniform column
 * indices drawn with replacement, duplicates merged into sqrt(count) * (first sign).  One CTA per sketch row regenerates the row and gathers
 * the <= vec_nnz rows of A it touches.  state <- S.next_state.  rlb200_fill_sparse_*_dev exports the same operator (major_axis = RLB200_AXIS_LONG;
 * its size query then returns the upper bound vec_nnz * #vectors, the call the exact count after merging).  Single shard. */
RLB200_API int rlb200_sketch_sparse_left_laso_f64_dev(rlb200_ctx* ctx, int64_t S_rows, int64_t S_cols, int64_t vec_nnz, int64_t d, int64_t n,
                                           int64_t m, double alpha, int64_t ro_s, int64_t co_s, const double* A_dev, int64_t lda, double beta,
                                           double* B_dev, int64_t ldb, uint32_t state[6]);
RLB200_API int rlb200_sketch_sparse_left_laso_f32_dev(rlb200_ctx* ctx, int64_t S_rows, int64_t S_cols, int64_t vec_nnz, int64_t d, int64_t n,
                                           int64_t m, float alpha, int64_t ro_s, int64_t co_s, const float* A_dev, int64_t lda, float beta,
                                           float* B_dev, int64_t ldb, uint32_t state[6]);

/* ---- a5: sketch_general with a DenseSkOp, left (lskge3, skge.hh:155-203) and right (rskge3, skge.hh:308-356), ColMajor, NoTrans.
 * S = DenseSkOp(DenseDist(S_rows, S_cols, family, major_axis), state); it is never materialised as a whole: panels are regenerated
 * from the Philox state into an L2-resident ring buffer and consumed by the tensor-pipe GEMM.
 *   left :  B(d x n) = alpha * S[ro_s:ro_s+d, co_s:co_s+m] * A(m x n) + beta * B
 *   right:  B(m x d) = alpha * A(m x n) * S[ro_s:ro_s+n, co_s:co_s+d] + beta * B
 * state <- S.next_state. */
RLB200_API int rlb200_sketch_dense_left_f64_dev(rlb200_ctx* ctx, int64_t S_rows, int64_t S_cols, int family, int major_axis, int64_t d,
                                     int64_t n, int64_t m, double alpha, int64_t ro_s, int64_t co_s, const double* A_dev, int64_t lda,
                                     double beta, double* B_dev, int64_t ldb, uint32_t state[6]);
RLB200_API int rlb200_sketch_dense_left_f32_dev(rlb200_ctx* ctx, int64_t S_rows, int64_t S_cols, int family, int major_axis, int64_t d,
                                     int64_t n, int64_t m, float alpha, int64_t ro_s, int64_t co_s, const float* A_dev, int64_t lda,
                                     float beta, float* B_dev, int64_t ldb, uint32_t state[6]);
RLB200_API int rlb200_sketch_dense_right_f64_dev(rlb200_ctx* ctx, int64_t S_rows, int64_t S_cols, int family, int major_axis, int64_t m,
                                      int64_t d, int64_t n, double alpha, const double* A_dev, int64_t lda, int64_t ro_s, int64_t co_s,
                                      double beta, double* B_dev, int64_t ldb, uint32_t state[6]);
RLB200_API int rlb200_sketch_dense_right_f32_dev(rlb200_ctx* ctx, int64_t S_rows, int64_t S_cols, int family, int major_axis, int64_t m,
                                      int64_t d, int64_t n, float alpha, const float* A_dev, int64_t lda, int64_t ro_s, int64_t co_s,
                                      float beta, float* B_dev, int64_t ldb, uint32_t state[6]);

/* sketch_general with every layout and transposition flag, dense operators (RandBLAS/RandBLAS/skge.hh:859-905 left -> lskge3 :100-203;
 * :1031-1076 right -> rskge3 :253-356).  layout: RLB200_LAYOUT_COLMAJOR | RLB200_LAYOUT_ROWMAJOR; opS / opA: 0 NoTrans, 1 Trans.
 *   left :  B(d x n) = alpha * op(submat(S))(d x m) * op(A)(m x n) + beta * B      submat(S) = the (d x m | m x d) block of S at (ro_s, co_s)
 *   right:  B(m x d) = alpha * op(A)(m x n) * op(submat(S))(n x d) + beta * B
 * lda / ldb follow `layout` as in the reference (:136-143, :289-296).  ColMajor with opA = NoTrans is the direct form; a data matrix that
 * arrives transposed is transposed once into device scratch (m * n extra elements), a RowMajor result is formed in scratch and written
 * through a transposing axpby.  state <- S.next_state.  Row-sharded contexts: the direct form only. */
RLB200_API int rlb200_sketch_general_dense_left_f64_dev(rlb200_ctx* ctx, int layout, int opS, int opA, int64_t d, int64_t n, int64_t m, double alpha,
                                             int64_t S_rows, int64_t S_cols, int family, int major_axis, int64_t ro_s, int64_t co_s,
                                             const double* A_dev, int64_t lda, double beta, double* B_dev, int64_t ldb, uint32_t state[6]);
RLB200_API int rlb200_sketch_general_dense_left_f32_dev(rlb200_ctx* ctx, int layout, int opS, int opA, int64_t d, int64_t n, int64_t m, float alpha,
                                             int64_t S_rows, int64_t S_cols, int family, int major_axis, int64_t ro_s, int64_t co_s,
                                             const float* A_dev, int64_t lda, float beta, float* B_dev, int64_t ldb, uint32_t state[6]);
RLB200_API int rlb200_sketch_general_dense_right_f64_dev(rlb200_ctx* ctx, int layout, int opA, int opS, int64_t m, int64_t d, int64_t n, double alpha,
                                              const double* A_dev, int64_t lda, int64_t S_rows, int64_t S_cols, int family, int major_axis,
                                              int64_t ro_s, int64_t co_s, double beta, double* B_dev, int64_t ldb, uint32_t state[6]);
RLB200_API int rlb200_sketch_general_dense_right_f32_dev(rlb200_ctx* ctx, int layout, int opA, int opS, int64_t m, int64_t d, int64_t n, float alpha,
                                              const float* A_dev, int64_t lda, int64_t S_rows, int64_t S_cols, int family, int major_axis,
                                              int64_t ro_s, int64_t co_s, float beta, float* B_dev, int64_t ldb, uint32_t state[6]);

/* The same for short-axis SparseSkOps (skge.hh:907-960 left -> lskges :538-571; :1078-1131 right -> rskges :573-620): op(submat(S)) must be
 * a WIDE operator, i.e. opS = NoTrans with a wide S or opS = Trans with a tall S (left), the reverse for the right sketch - the shapes a
 * sketch has; the operator is regenerated on the device from the state (a tall short-axis operator is the transpose of the wide one with the
 * same seed).  The right sketch runs as the left sketch of the transposed problem, as in the reference. */
RLB200_API int rlb200_sketch_general_sparse_left_f64_dev(rlb200_ctx* ctx, int layout, int opS, int opA, int64_t d, int64_t n, int64_t m, double alpha,
                                              int64_t S_rows, int64_t S_cols, int64_t vec_nnz, int64_t ro_s, int64_t co_s, const double* A_dev,
                                              int64_t lda, double beta, double* B_dev, int64_t ldb, uint32_t state[6]);
RLB200_API int rlb200_sketch_general_sparse_left_f32_dev(rlb200_ctx* ctx, int layout, int opS, int opA, int64_t d, int64_t n, int64_t m, float alpha,
                                              int64_t S_rows, int64_t S_cols, int64_t vec_nnz, int64_t ro_s, int64_t co_s, const float* A_dev,
                                              int64_t lda, float beta, float* B_dev, int64_t ldb, uint32_t state[6]);
RLB200_API int rlb200_sketch_general_sparse_right_f64_dev(rlb200_ctx* ctx, int layout, int opA, int opS, int64_t m, int64_t d, int64_t n, double alpha,
                                               const double* A_dev, int64_t lda, int64_t S_rows, int64_t S_cols, int64_t vec_nnz, int64_t ro_s,
                                               int64_t co_s, double beta, double* B_dev, int64_t ldb, uint32_t state[6]);
RLB200_API int rlb200_sketch_general_sparse_right_f32_dev(rlb200_ctx* ctx, int layout, int opA, int opS, int64_t m, int64_t d, int64_t n, float alpha,
                                               const float* A_dev, int64_t lda, int64_t S_rows, int64_t S_cols, int64_t vec_nnz, int64_t ro_s,
                                               int64_t co_s, float beta, float* B_dev, int64_t ldb, uint32_t state[6]);

/* ---- blas::gemm as used on the path (ColMajor; rl_rs.hh:142,153,165; rl_rf.hh:123; rl_qb.hh:218;
 *      rl_rsvd.hh:148).  transa/transb: 0 = NoTrans, 1 = Trans.  Shapes the tall-skinny kernels cover:
 *      NN with the long dimension on m; TN with the long dimension on k (split-K, deterministic). */
RLB200_API int rlb200_gemm_f64_dev(rlb200_ctx* ctx, int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha,
                        const double* A_dev, int64_t lda, const double* B_dev, int64_t ldb, double beta, double* C_dev, int64_t ldc);
RLB200_API int rlb200_gemm_f32_dev(rlb200_ctx* ctx, int transa, int transb, int64_t m, int64_t n, int64_t k, float alpha,
                        const float* A_dev, int64_t lda, const float* B_dev, int64_t ldb, float beta, float* C_dev, int64_t ldc);

/* ---- the same blas::gemm call sites on the tcgen05 tensor cores: operands are split on the fly into exact balanced base-256
 *      int8 digit slices (S digits keep 8S-2 bits below each row's / column-chunk's largest magnitude), multiplied with
 *      tcgen05.mma.kind::i8 with exact int32 accumulation in TMEM and recombined in fp64 (Ozaki scheme).  Shapes: NN (tall A) and
 *      TN (long contraction), as rlb200_gemm_f64_dev. */
RLB200_API int rlb200_gemm_f64_i8_dev(rlb200_ctx* ctx, int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha,
                           const double* A_dev, int64_t lda, const double* B_dev, int64_t ldb, double beta, double* C_dev, int64_t ldc);
RLB200_API int rlb200_gemm_f32_i8_dev(rlb200_ctx* ctx, int transa, int transb, int64_t m, int64_t n, int64_t k, float alpha,
                           const float* A_dev, int64_t lda, const float* B_dev, int64_t ldb, float beta, float* C_dev, int64_t ldc);
/* Engine used by the drivers (RS/RF/QB/RSVD, CQRRPT, BQRRP) for their tall products (m >= 16384 rows; shorter ones always use the
 * fp64 pipe): RLB200_FP64_I8SLICES (default) or RLB200_FP64_DMMA. */
enum { RLB200_FP64_DMMA = 0, RLB200_FP64_I8SLICES = 1 };
RLB200_API int rlb200_set_fp64_engine(rlb200_ctx* ctx, int engine);
/* Fused engine (default on): tall products whose shapes allow it (second dimension >= 96) produce the digits of the tall operand inside
 * the tensor-core kernel (one fp64 read of the data matrix per pass, no digit round trip through HBM); off = always stage the digits. */
RLB200_API int rlb200_set_i8_fused(rlb200_ctx* ctx, int on);
/* Digits per value of the int8-slice engine: 0 = default (6 for fp64 storage: 46 bits; 4 for fp32: 30 bits), else 3..7. */
RLB200_API int rlb200_set_i8_digits(rlb200_ctx* ctx, int digits);

/* ---- a8/a9: Stabilization<T>::call(m, k, A) (rl_orth.hh:13-23): CholQRQ :68-98, HQRQ :144-164,
 *      PLUL :211-230.  In place on the m x k column-major A_dev (lda = m).
 *      Returns 0, or 1 exactly where the reference does (potrf failure -> chol_fail; cond check). */
RLB200_API int rlb200_stab_f64_dev(rlb200_ctx* ctx, int kind, int64_t m, int64_t k, double* A_dev, int cond_check, int* chol_fail);
RLB200_API int rlb200_stab_f32_dev(rlb200_ctx* ctx, int kind, int64_t m, int64_t k, float* A_dev, int cond_check, int* chol_fail);

/* ---- a10: RS<T,RNG>::call(m, n, A, k, Omega, state) (rl_rs.hh:116-178) ----------------------
 * Omega_dev: n x k.  work_dev: m x k scratch (the reference's Omega_1), may be NULL when p == 0.
 * Returns 0 / 1 (stabiliser failure). */
RLB200_API int rlb200_rs_f64_dev(rlb200_ctx* ctx, int64_t m, int64_t n, const double* A_dev, int64_t k, double* Omega_dev,
                      double* work_dev, uint32_t state[6], const rlb200_stack_opts* opts);
RLB200_API int rlb200_rs_f32_dev(rlb200_ctx* ctx, int64_t m, int64_t n, const float* A_dev, int64_t k, float* Omega_dev,
                      float* work_dev, uint32_t state[6], const rlb200_stack_opts* opts);

/* ---- a11: RF<T,RNG>::call(m, n, A, k, Q, state) (rl_rf.hh:106-137) --------------------------
 * Q_dev: m x k (also used as RS scratch before being overwritten).  Returns 0 / 1 / 2. */
RLB200_API int rlb200_rf_f64_dev(rlb200_ctx* ctx, int64_t m, int64_t n, const double* A_dev, int64_t k, double* Q_dev,
                      uint32_t state[6], const rlb200_stack_opts* opts);
RLB200_API int rlb200_rf_f32_dev(rlb200_ctx* ctx, int64_t m, int64_t n, const float* A_dev, int64_t k, float* Q_dev,
                      uint32_t state[6], const rlb200_stack_opts* opts);

/* ---- a12: QB<T,RNG>::call(m, n, A, k, block_sz, tol, Q, BT, state) (rl_qb.hh:133-268) -------
 * Q_dev: m x k, BT_dev: n x k caller-allocated for the requested k; *k is in/out as in the reference.
 * With block_sz == k (single block) A_dev is only read; with block_sz < k the reference deflates a
 * COPY of A (rl_qb.hh:162,171,260): pass Acpy_dev (m x n scratch) or NULL to deflate A_dev in place.
 * Returns the reference's codes 0/2/3/4/5/6. */
RLB200_API int rlb200_qb_f64_dev(rlb200_ctx* ctx, int64_t m, int64_t n, double* A_dev, int64_t* k, int64_t block_sz, double tol,
                      double* Q_dev, double* BT_dev, double* Acpy_dev, uint32_t state[6], const rlb200_stack_opts* opts);
RLB200_API int rlb200_qb_f32_dev(rlb200_ctx* ctx, int64_t m, int64_t n, float* A_dev, int64_t* k, int64_t block_sz, float tol,
                      float* Q_dev, float* BT_dev, float* Acpy_dev, uint32_t state[6], const rlb200_stack_opts* opts);

/* ---- a13: RSVD<T,RNG>::call(m, n, A, k, tol, U, S, V, state) (rl_rsvd.hh:113-154) -----------
 * U_dev: m x k, S_dev: k, V_dev: n x k for the requested k; *k in/out.  U is formed in the Q buffer
 * (U_dev doubles as Q), so no second m x k allocation exists.  Returns 0 (the reference ignores QB's code);
 * qb_code (optional) receives QB's code. */
RLB200_API int rlb200_rsvd_f64_dev(rlb200_ctx* ctx, int64_t m, int64_t n, double* A_dev, int64_t* k, double tol, double* U_dev,
                        double* S_dev, double* V_dev, double* Acpy_dev, uint32_t state[6], const rlb200_stack_opts* opts, int* qb_code);
RLB200_API int rlb200_rsvd_f32_dev(rlb200_ctx* ctx, int64_t m, int64_t n, float* A_dev, int64_t* k, float tol, float* U_dev,
                        float* S_dev, float* V_dev, float* Acpy_dev, uint32_t state[6], const rlb200_stack_opts* opts, int* qb_code);

/* Host-pointer form of the same call: A, U, S, V are HOST buffers (U: m x k, S: k, V: n x k for the
 * requested k, caller-allocated); copies A to the device, runs rlb200_rsvd_*_dev, copies results back. */
RLB200_API int rlb200_rsvd_f64_host(rlb200_ctx* ctx, int64_t m, int64_t n, const double* A, int64_t* k, double tol, double* U,
                         double* S, double* V, uint32_t state[6], const rlb200_stack_opts* opts, int* qb_code);
RLB200_API int rlb200_rsvd_f32_host(rlb200_ctx* ctx, int64_t m, int64_t n, const float* A, int64_t* k, float tol, float* U,
                         float* S, float* V, uint32_t state[6], const rlb200_stack_opts* opts, int* qb_code);

/* ---- a14: CQRRPT<T,RNG>::call(m, n, A, lda, R, ldr, J, d_factor, state) (RandLAPACK/drivers/rl_cqrrpt.hh:146-391) with the
 *      default QRCP (geqp3, :56-66,247).  `eps` and `nnz` are the object's public fields (ctor argument `ep`; SASO nnz, default 2).
 * A_dev (m x n, lda >= m) is overwritten by Q (its first *rank columns); R_dev (ldr >= n, n columns) receives the rank x n
 * upper-trapezoidal factor (only the entries the reference writes are written); J_dev: n 1-based GEQP3-style pivots;
 * *rank <- this->rank.  A[:, J] = Q R.  Returns 0, or 1 exactly where the reference does (:300-305).
 * Row-sharded contexts: A_dev holds this rank's row block; the sketch and the Gram matrix are allreduced, R/J/rank are replicated. */
RLB200_API int rlb200_cqrrpt_f64_dev(rlb200_ctx* ctx, int64_t m, int64_t n, double* A_dev, int64_t lda, double* R_dev, int64_t ldr,
                          int64_t* J_dev, double d_factor, double eps, int64_t nnz, int64_t* rank, uint32_t state[6]);
RLB200_API int rlb200_cqrrpt_f32_dev(rlb200_ctx* ctx, int64_t m, int64_t n, float* A_dev, int64_t lda, float* R_dev, int64_t ldr,
                          int64_t* J_dev, float d_factor, float eps, int64_t nnz, int64_t* rank, uint32_t state[6]);
/* Host-pointer form (the reference's own calling convention): A, R, J are HOST buffers. */
RLB200_API int rlb200_cqrrpt_f64_host(rlb200_ctx* ctx, int64_t m, int64_t n, double* A, int64_t lda, double* R, int64_t ldr, int64_t* J,
                           double d_factor, double eps, int64_t nnz, int64_t* rank, uint32_t state[6]);
RLB200_API int rlb200_cqrrpt_f32_host(rlb200_ctx* ctx, int64_t m, int64_t n, float* A, int64_t lda, float* R, int64_t ldr, int64_t* J,
                           float d_factor, float eps, int64_t nnz, int64_t* rank, uint32_t state[6]);

/* ---- f1: CQRRT<T,RNG>::call(m, n, A, lda, R, ldr, d_factor, state) (RandLAPACK/drivers/rl_cqrrt.hh:20-37 base, :91-297 body): unpivoted
 *      sketched Cholesky QR.  nnz, orthogonalization, compute_Q are the object's public fields (ctor defaults 2, false, true, :46-50).
 * On exit A holds Q (m x n; A R_sk^-1 only, when compute_Q == 0), the upper triangle of R the n x n factor (R_chol itself when
 * orthogonalization != 0); state <- S.next_state.  Returns the reference's codes: 0, or 1 when the sketch's R has a zero diagonal entry
 * (:173-177) or the Cholesky factorization fails (:194-198).  Row-shardable like CQRRPT (sketch and Gram matrix are sum-allreduced). */
RLB200_API int rlb200_cqrrt_f64_dev(rlb200_ctx* ctx, int64_t m, int64_t n, double* A_dev, int64_t lda, double* R_dev, int64_t ldr, double d_factor,
                         int64_t nnz, int orthogonalization, int compute_Q, uint32_t state[6]);
RLB200_API int rlb200_cqrrt_f32_dev(rlb200_ctx* ctx, int64_t m, int64_t n, float* A_dev, int64_t lda, float* R_dev, int64_t ldr, float d_factor,
                         int64_t nnz, int orthogonalization, int compute_Q, uint32_t state[6]);
RLB200_API int rlb200_cqrrt_f64_host(rlb200_ctx* ctx, int64_t m, int64_t n, double* A, int64_t lda, double* R, int64_t ldr, double d_factor,
                          int64_t nnz, int orthogonalization, int compute_Q, uint32_t state[6]);
RLB200_API int rlb200_cqrrt_f32_host(rlb200_ctx* ctx, int64_t m, int64_t n, float* A, int64_t lda, float* R, int64_t ldr, float d_factor,
                          int64_t nnz, int orthogonalization, int compute_Q, uint32_t state[6]);

/* ---- a15: BQRRP<T,RNG>::call(m, n, A, lda, d_factor, tau, J, state) (RandLAPACK/drivers/rl_bqrrp.hh:154-665), and the
 *      device-pointer convention of BQRRP_GPU::call (rl_bqrrp_gpu.hh:152-942).  block_size, qrcp_wide, qr_tall are the object's
 *      fields: block_size = ctor's b_sz; qrcp_wide: 0 = luqr (default), 1 = geqp3; qr_tall: 0 = geqrf (default), 1 = cholqr
 *      (CholQR + Householder reconstruction; the choice BQRRP_GPU makes), 2 = geqrt.  tol = eps (ctor default).
 * On exit A_dev is GEQP3-formatted (R in the upper triangle, Householder vectors below), tau_dev (n; entries written as the
 * reference writes them), J_dev (n 1-based pivots), *rank <- this->rank.  The d x n Gaussian sketch is formed internally from `state`
 * exactly as the CPU reference does (:309-312) with S regenerated on chip; state <- the reference's advanced state.
 * Not row-shardable (RLB200_ERR_UNSUPPORTED on a sharded context): replicas only. */
enum { RLB200_QRCP_LUQR = 0, RLB200_QRCP_GEQP3 = 1 };
enum { RLB200_QRTALL_GEQRF = 0, RLB200_QRTALL_CHOLQR = 1, RLB200_QRTALL_GEQRT = 2 };
RLB200_API int rlb200_bqrrp_f64_dev(rlb200_ctx* ctx, int64_t m, int64_t n, double* A_dev, int64_t lda, double d_factor, int64_t block_size,
                         int qrcp_wide, int qr_tall, double* tau_dev, int64_t* J_dev, int64_t* rank, uint32_t state[6]);
RLB200_API int rlb200_bqrrp_f32_dev(rlb200_ctx* ctx, int64_t m, int64_t n, float* A_dev, int64_t lda, float d_factor, int64_t block_size,
                         int qrcp_wide, int qr_tall, float* tau_dev, int64_t* J_dev, int64_t* rank, uint32_t state[6]);
/* BQRRP_GPU_alg<T,RNG>::call(m, n, A, lda, A_sk, d, tau, J) (RandLAPACK/drivers/rl_bqrrp_gpu.hh:27-43, 122-133; body :152-942): every
 * pointer is a DEVICE pointer and the d x n sketch A_sk (leading dimension d, d >= block_size) is an INPUT - formed by the caller, e.g. as
 * S * A with S = fill_dense(DenseDist(d, m)) (test/drivers/test_bqrrp_gpu.cu:91-103) - and is overwritten.  qrcp_wide is LUQR (the only
 * choice the reference's GPU driver offers, :54); qr_tall: RLB200_QRTALL_GEQRF (its ctor default, :86) or RLB200_QRTALL_CHOLQR.
 * Outputs as rlb200_bqrrp_*_dev. */
RLB200_API int rlb200_bqrrp_f64_dev_sk(rlb200_ctx* ctx, int64_t m, int64_t n, double* A_dev, int64_t lda, double* A_sk_dev, int64_t d,
                            int64_t block_size, int qr_tall, double* tau_dev, int64_t* J_dev, int64_t* rank);
RLB200_API int rlb200_bqrrp_f32_dev_sk(rlb200_ctx* ctx, int64_t m, int64_t n, float* A_dev, int64_t lda, float* A_sk_dev, int64_t d,
                            int64_t block_size, int qr_tall, float* tau_dev, int64_t* J_dev, int64_t* rank);
/* Host-pointer form (the CPU reference's calling convention): A, tau, J are HOST buffers. */
RLB200_API int rlb200_bqrrp_f64_host(rlb200_ctx* ctx, int64_t m, int64_t n, double* A, int64_t lda, double d_factor, int64_t block_size,
                          int qrcp_wide, int qr_tall, double* tau, int64_t* J, int64_t* rank, uint32_t state[6]);
RLB200_API int rlb200_bqrrp_f32_host(rlb200_ctx* ctx, int64_t m, int64_t n, float* A, int64_t lda, float d_factor, int64_t block_size,
                          int qrcp_wide, int qr_tall, float* tau, int64_t* J, int64_t* rank, uint32_t state[6]);

/* ---- f2: hqrrp(m, n, A, lda, jpvt, tau, nb_alg, pp, panel_pivoting, qr_type, state, timing) (RandLAPACK/drivers/rl_hqrrp.hh:811-1196):
 *      Householder QR with randomized pivoting.  nb_alg = block size, pp = oversampling (the sketch has nb_alg + pp rows),
 *      panel_pivoting != 0: every panel is factored by norm-downdating QRCP (:556-775); otherwise qr_type picks the panel QR: 0 the unblocked
 *      Householder loop, 1 geqrf (:464-502), 2 CholQR + Householder reconstruction (:505-553; needs m - j >= nb_alg in every block).
 * On exit A is GEQP3-formatted, tau holds min(m, n) scalars, J n 1-based pivots (J is NOT written when min(m, n) = 0, :886-888);
 * state <- fill_dense(DenseDist(nb_alg + pp, m, Uniform)).next_state (:928-929).  Returns 0 (the reference's only return value), or 1 when a
 * panel's Cholesky factorization fails under qr_type 2 (the reference continues with an unfactored panel there).  With
 * rlb200_set_phase_timing the nine leading entries of the reference's timing vector (:1140-1148) are recorded.  Replicas only (not row-shardable). */
RLB200_API int rlb200_hqrrp_f64_dev(rlb200_ctx* ctx, int64_t m, int64_t n, double* A_dev, int64_t lda, int64_t* J_dev, double* tau_dev,
                         int64_t nb_alg, int64_t pp, int panel_pivoting, int qr_type, uint32_t state[6]);
RLB200_API int rlb200_hqrrp_f32_dev(rlb200_ctx* ctx, int64_t m, int64_t n, float* A_dev, int64_t lda, int64_t* J_dev, float* tau_dev,
                         int64_t nb_alg, int64_t pp, int panel_pivoting, int qr_type, uint32_t state[6]);
/* Host-pointer form (the reference's calling convention): A, J, tau are HOST buffers. */
RLB200_API int rlb200_hqrrp_f64_host(rlb200_ctx* ctx, int64_t m, int64_t n, double* A, int64_t lda, int64_t* J, double* tau,
                          int64_t nb_alg, int64_t pp, int panel_pivoting, int qr_type, uint32_t state[6]);
RLB200_API int rlb200_hqrrp_f32_host(rlb200_ctx* ctx, int64_t m, int64_t n, float* A, int64_t lda, int64_t* J, float* tau,
                          int64_t nb_alg, int64_t pp, int panel_pivoting, int qr_type, uint32_t state[6]);

/* ---- lapack::geqp3 / geqrf of a small (L2-resident) d x n matrix, as used on the sketch (rl_cqrrpt.hh:247, rl_bqrrp.hh:336,356).
 * pivot != 0: J_dev receives n 1-based pivots (all columns free on entry, i.e. LAPACK's jpvt = 0 convention of the call sites). */
RLB200_API int rlb200_qr_small_f64_dev(rlb200_ctx* ctx, int pivot, int64_t d, int64_t n, double* A_dev, int64_t lda, int64_t* J_dev,
                            double* tau_dev);
RLB200_API int rlb200_qr_small_f32_dev(rlb200_ctx* ctx, int pivot, int64_t d, int64_t n, float* A_dev, int64_t lda, int64_t* J_dev,
                            float* tau_dev);
/* ---- util::col_swap(m, n, k, A, lda, idx) = lapack::lapmt(forward) (RandLAPACK/misc/rl_util.hh:151-165): column i of A <- old
 *      column idx[i]-1; idx_host: n 1-based entries on the HOST (left unchanged). */
RLB200_API int rlb200_col_swap_f64_dev(rlb200_ctx* ctx, int64_t m, int64_t n, double* A_dev, int64_t lda, const int64_t* idx_host);
RLB200_API int rlb200_col_swap_f32_dev(rlb200_ctx* ctx, int64_t m, int64_t n, float* A_dev, int64_t lda, const int64_t* idx_host);

/* ---- lapack::gesdd(SomeVec) of a tall n x k matrix as used at rl_rsvd.hh:146 ----------------
 * B_dev (n x k, ld n) is overwritten by its left singular vectors (n x k), S_dev by the singular
 * values (descending), W_dev (k x k, ld k) by the RIGHT singular vectors as columns (so B_in = B_out diag(S) W^T). */
RLB200_API int rlb200_svd_tall_f64_dev(rlb200_ctx* ctx, int64_t n, int64_t k, double* B_dev, double* S_dev, double* W_dev);
RLB200_API int rlb200_svd_tall_f32_dev(rlb200_ctx* ctx, int64_t n, int64_t k, float* B_dev, float* S_dev, float* W_dev);

/* ---- f3: SYPS / SYRF / REVD2 on an explicit symmetric matrix (RandLAPACK/comps/rl_syps.hh:21-143, comps/rl_syrf.hh:21-118,
 *      drivers/rl_revd2.hh:75-246).  A_dev: m x m column-major, lda >= m; only the `uplo` triangle is read (ExplicitSymLinOp,
 *      linops/rl_sym_linops.hh; the other triangle may hold anything, NaN included).  Replicas only (not row-shardable).
 * SYPS::call(uplo, m, A, lda, k, state, skop_buff, work_buff): skop_dev (m x k) <- the power-sketched operator, work_dev: m x k scratch;
 *   state <- the state after the DenseDist(m, k) sample.  The stabiliser is Householder QR (geqrf + ungqr), k <= 256.
 * SYRF::call(uplo, m, A, k, Q, state, work_buff): Q_dev (m x k) <- orth(A * syps(A)); work_dev: m x k scratch.  Returns 0, or 2 when the
 *   orthogonaliser fails (the reference throws).
 * REVD2::call(uplo, m, A, k, tol, V, eigvals, state): V_dev (m x k_cap), eigvals_dev (k_cap); *k in/out: the rank estimate doubles (capped at m)
 *   until the error estimate <= 5 max(tol, nu) (:225-231).  Returns 0; 1 = Cholesky of the regularised core failed, 2 = orthogonaliser
 *   failed (both std::runtime_error in the reference); 3 = the next k would exceed k_cap (outputs hold the last iterate; the reference
 *   would have grown its std::vectors).  err_est (optional) <- the last error estimate. */
RLB200_API int rlb200_syps_f64_dev(rlb200_ctx* ctx, int uplo, int64_t m, const double* A_dev, int64_t lda, int64_t k, int64_t passes,
                        int64_t passes_per_stab, double* skop_dev, double* work_dev, uint32_t state[6]);
RLB200_API int rlb200_syps_f32_dev(rlb200_ctx* ctx, int uplo, int64_t m, const float* A_dev, int64_t lda, int64_t k, int64_t passes,
                        int64_t passes_per_stab, float* skop_dev, float* work_dev, uint32_t state[6]);
RLB200_API int rlb200_syrf_f64_dev(rlb200_ctx* ctx, int uplo, int64_t m, const double* A_dev, int64_t lda, int64_t k, double* Q_dev,
                        double* work_dev, uint32_t state[6], const rlb200_revd2_opts* opts);
RLB200_API int rlb200_syrf_f32_dev(rlb200_ctx* ctx, int uplo, int64_t m, const float* A_dev, int64_t lda, int64_t k, float* Q_dev,
                        float* work_dev, uint32_t state[6], const rlb200_revd2_opts* opts);
RLB200_API int rlb200_revd2_f64_dev(rlb200_ctx* ctx, int uplo, int64_t m, const double* A_dev, int64_t lda, int64_t* k, int64_t k_cap, double tol,
                         double* V_dev, double* eigvals_dev, uint32_t state[6], const rlb200_revd2_opts* opts, double* err_est);
RLB200_API int rlb200_revd2_f32_dev(rlb200_ctx* ctx, int uplo, int64_t m, const float* A_dev, int64_t lda, int64_t* k, int64_t k_cap, float tol,
                         float* V_dev, float* eigvals_dev, uint32_t state[6], const rlb200_revd2_opts* opts, float* err_est);
/* Host-pointer form (the reference's calling convention): A, V (m x k_cap), eigvals (k_cap) are HOST buffers. */
RLB200_API int rlb200_revd2_f64_host(rlb200_ctx* ctx, int uplo, int64_t m, const double* A, int64_t lda, int64_t* k, int64_t k_cap, double tol,
                          double* V, double* eigvals, uint32_t state[6], const rlb200_revd2_opts* opts, double* err_est);
RLB200_API int rlb200_revd2_f32_host(rlb200_ctx* ctx, int uplo, int64_t m, const float* A, int64_t lda, int64_t* k, int64_t k_cap, float tol,
                          float* V, float* eigvals, uint32_t state[6], const rlb200_revd2_opts* opts, float* err_est);

#ifdef __cplusplus
}
#endif
#endif /* RLB200_H */
