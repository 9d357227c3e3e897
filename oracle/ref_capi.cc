// oracle/_ref/librl_ref.so — the UNMODIFIED reference headers, compiled where they lie under
// /root/reference, behind a flat C API.  TEST INFRASTRUCTURE ONLY: only tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library.
//
// Nothing from the reference is copied: this file only #includes its headers
// (-I/root/reference, -I/root/reference/RandBLAS) against the shim BLAS++/LAPACK++/Random123
// headers in oracle/shim/ (ours), and forwards.  Every entry point names the reference symbol
// it calls.  The same signatures (prefix rlo_) are exported by the independent restatement in
// oracle/rl_oracle.c + oracle/rl_oracle.py, and (prefix rlb200_, device pointers) by the product.
#include <RandBLAS.hh>
#include "RandLAPACK/rl_blaspp.hh"
#include "RandLAPACK/rl_lapackpp.hh"
#include "RandLAPACK/rl_exceptions.hh"
#include "RandLAPACK/misc/rl_util.hh"
#include "RandLAPACK/comps/rl_orth.hh"
#include "RandLAPACK/comps/rl_rs.hh"
#include "RandLAPACK/comps/rl_rf.hh"
#include "RandLAPACK/comps/rl_qb.hh"
#include "RandLAPACK/drivers/rl_rsvd.hh"
#include "RandLAPACK/drivers/rl_cqrrpt.hh"
#include "RandLAPACK/drivers/rl_cqrrt.hh"
#include "RandLAPACK/drivers/rl_bqrrp.hh"
#include "RandLAPACK/comps/rl_syps.hh"
#include "RandLAPACK/comps/rl_syrf.hh"
#include "RandLAPACK/drivers/rl_revd2.hh"
#include "RandLAPACK/testing/rl_gen.hh"

#include <cstring>
#include <memory>
#include <chrono>

#include "oracle_capi.h"

using RNG = r123::Philox4x32;
using State = RandBLAS::RNGState<RNG>;

static State load_state(const uint32_t s[6]) {
    State st;
    for (int i = 0; i < 4; ++i) st.counter.v[i] = s[i];
    for (int i = 0; i < 2; ++i) st.key.v[i] = s[4 + i];
    return st;
}
static void store_state(const State& st, uint32_t s[6]) {
    for (int i = 0; i < 4; ++i) s[i] = st.counter.v[i];
    for (int i = 0; i < 2; ++i) s[4 + i] = st.key.v[i];
}

template <typename T>
static std::unique_ptr<RandLAPACK::Stabilization<T>> make_stab(int kind, bool cond_check) {
    switch (kind) {
        case RL_STAB_PLUL:    return std::make_unique<RandLAPACK::PLUL<T>>(cond_check, false);
        case RL_STAB_CHOLQRQ: return std::make_unique<RandLAPACK::CholQRQ<T>>(cond_check, false);
        case RL_STAB_HQRQ:    return std::make_unique<RandLAPACK::HQRQ<T>>(cond_check, false);
    }
    throw std::runtime_error("bad stabiliser kind");
}

// the canonical stack of test/drivers/test_rsvd.cc:68-93
template <typename T>
struct Stack {
    std::unique_ptr<RandLAPACK::Stabilization<T>> stab, orth_rf, orth_qb;
    RandLAPACK::RS<T, RNG> rs;
    RandLAPACK::RF<T, RNG> rf;
    RandLAPACK::QB<T, RNG> qb;
    RandLAPACK::RSVD<T, RNG> rsvd;
    explicit Stack(const rl_stack_opts& o)
        : stab(make_stab<T>(o.stab, o.cond_check)), orth_rf(make_stab<T>(o.orth_rf, o.cond_check)),
          orth_qb(make_stab<T>(o.orth_qb, o.cond_check)),
          rs(*stab, o.passes_over_data, o.passes_per_stab, false, o.cond_check),
          rf(rs, *orth_rf, false, o.cond_check), qb(rf, *orth_qb, false, o.orth_check), rsvd(qb, o.block_sz) {}
};

#define RL_TRY try {
#define RL_CATCH } catch (const std::exception& e) { std::snprintf(g_err, sizeof g_err, "%s", e.what()); return RL_ERR_EXCEPTION; }
static thread_local char g_err[512];

template <typename T>
static int fill_dense_impl(int64_t n_rows, int64_t n_cols, int family, int major_axis, int layout, int64_t sub_rows,
                           int64_t sub_cols, int64_t ro, int64_t co, T* buff, uint32_t state[6]) {
    RL_TRY
    RandBLAS::DenseDist D(n_rows, n_cols, family == RL_FAMILY_UNIFORM ? RandBLAS::ScalarDist::Uniform : RandBLAS::ScalarDist::Gaussian,
                          major_axis == RL_AXIS_SHORT ? RandBLAS::Axis::Short : RandBLAS::Axis::Long);
    blas::Layout lay = layout == RL_LAYOUT_NATURAL ? D.natural_layout
                       : (layout == RL_LAYOUT_ROWMAJOR ? blas::Layout::RowMajor : blas::Layout::ColMajor);
    State st = load_state(state);
    // RandBLAS/RandBLAS/dense_skops.hh:560-603 (fill_dense_unpacked); :620-623 is the full-matrix special case
    State nxt = RandBLAS::fill_dense_unpacked(lay, D, sub_rows, sub_cols, ro, co, buff, st);
    store_state(nxt, state);
    return 0;
    RL_CATCH
}

template <typename T>
static int rs_impl(int64_t m, int64_t n, const T* A, int64_t k, T* Omega, uint32_t state[6], const rl_stack_opts* o) {
    RL_TRY
    Stack<T> s(*o);
    State st = load_state(state);
    int rc = s.rs.call(m, n, A, k, Omega, st);   // RandLAPACK/comps/rl_rs.hh:116-178
    store_state(st, state);
    return rc;
    RL_CATCH
}

template <typename T>
static int rf_impl(int64_t m, int64_t n, const T* A, int64_t k, T* Q, uint32_t state[6], const rl_stack_opts* o) {
    RL_TRY
    Stack<T> s(*o);
    State st = load_state(state);
    int rc = s.rf.call(m, n, A, k, Q, st);       // RandLAPACK/comps/rl_rf.hh:106-137
    store_state(st, state);
    return rc;
    RL_CATCH
}

template <typename T>
static int qb_impl(int64_t m, int64_t n, T* A, int64_t* k, int64_t b_sz, T tol, T* Q_out, T* BT_out, uint32_t state[6],
                   const rl_stack_opts* o) {
    RL_TRY
    Stack<T> s(*o);
    State st = load_state(state);
    T *Q = nullptr, *BT = nullptr;
    int64_t k_io = *k;
    int rc = s.qb.call(m, n, A, k_io, b_sz, tol, Q, BT, st);   // RandLAPACK/comps/rl_qb.hh:133-268
    std::memcpy(Q_out, Q, sizeof(T) * m * k_io);
    std::memcpy(BT_out, BT, sizeof(T) * n * k_io);
    free(Q); free(BT);
    *k = k_io;
    store_state(st, state);
    return rc;
    RL_CATCH
}

template <typename T>
static int rsvd_impl(int64_t m, int64_t n, T* A, int64_t* k, T tol, T* U_out, T* S_out, T* V_out, uint32_t state[6],
                     const rl_stack_opts* o) {
    RL_TRY
    Stack<T> s(*o);
    State st = load_state(state);
    T *U = nullptr, *S = nullptr, *V = nullptr;
    int64_t k_io = *k;
    int rc = s.rsvd.call(m, n, A, k_io, tol, U, S, V, st);     // RandLAPACK/drivers/rl_rsvd.hh:113-154
    std::memcpy(U_out, U, sizeof(T) * m * k_io);
    std::memcpy(S_out, S, sizeof(T) * k_io);
    std::memcpy(V_out, V, sizeof(T) * n * k_io);
    free(U); free(S); free(V);
    *k = k_io;
    store_state(st, state);
    return rc;
    RL_CATCH
}

template <typename T>
static int stab_impl(int kind, int64_t m, int64_t k, T* A, int cond_check) {
    RL_TRY
    auto st = make_stab<T>(kind, cond_check);   // RandLAPACK/comps/rl_orth.hh:68-98,144-164,211-230
    return st->call(m, k, A);
    RL_CATCH
}

template <typename T>
static int mat_gen_impl(int type, int64_t m, int64_t n, int64_t rank, T cond, T exponent, T scaling, T* A, uint32_t state[6]) {
    RL_TRY
    RandLAPACK::gen::mat_gen_info<T> info(m, n, (RandLAPACK::gen::mat_type)type);
    info.rank = rank; info.cond_num = cond; info.exponent = exponent; info.scaling = scaling;
    State st = load_state(state);
    RandLAPACK::gen::mat_gen(info, A, st);      // RandLAPACK/testing/rl_gen.hh:712-773
    store_state(st, state);
    return 0;
    RL_CATCH
}

template <typename T>
static int fill_sparse_impl(int64_t n_rows, int64_t n_cols, int64_t vec_nnz, int major_axis, int64_t sub_rows, int64_t sub_cols,
                            int64_t ro, int64_t co, int64_t* nnz, T* vals, int64_t* rows, int64_t* cols, uint32_t state[6]) {
    RL_TRY
    RandBLAS::SparseDist D(n_rows, n_cols, vec_nnz, major_axis == RL_AXIS_SHORT ? RandBLAS::Axis::Short : RandBLAS::Axis::Long);
    State st = load_state(state);
    // RandBLAS/RandBLAS/sparse_skops.hh:568-704
    State nxt = RandBLAS::fill_sparse_unpacked(D, sub_rows, sub_cols, ro, co, *nnz, vals, rows, cols, st);
    store_state(nxt, state);
    return 0;
    RL_CATCH
}

template <typename T>
static int sketch_sparse_left_impl(int64_t S_rows, int64_t S_cols, int64_t vec_nnz, int64_t d, int64_t n, int64_t m, T alpha, int64_t ro,
                                   int64_t co, const T* A, int64_t lda, T beta, T* B, int64_t ldb, uint32_t state[6]) {
    RL_TRY
    RandBLAS::SparseDist DS(S_rows, S_cols, vec_nnz);
    State st = load_state(state);
    RandBLAS::SparseSkOp<T, RNG> S(DS, st);      // as at RandLAPACK/drivers/rl_cqrrpt.hh:214-216
    store_state(S.next_state, state);
    RandBLAS::sketch_general(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::NoTrans, d, n, m, alpha, S, ro, co, A, lda, beta, B, ldb);
    return 0;
    RL_CATCH
}

template <typename T>
static int sketch_dense_impl(bool left, int64_t S_rows, int64_t S_cols, int family, int major_axis, int64_t d, int64_t n, int64_t m,
                             T alpha, int64_t ro, int64_t co, const T* A, int64_t lda, T beta, T* B, int64_t ldb, uint32_t state[6]) {
    RL_TRY
    RandBLAS::DenseDist D(S_rows, S_cols, family == RL_FAMILY_UNIFORM ? RandBLAS::ScalarDist::Uniform : RandBLAS::ScalarDist::Gaussian,
                          major_axis == RL_AXIS_SHORT ? RandBLAS::Axis::Short : RandBLAS::Axis::Long);
    State st = load_state(state);
    RandBLAS::DenseSkOp<T, RNG> S(D, st);
    store_state(S.next_state, state);
    if (left)   // RandBLAS/RandBLAS/skge.hh:883-905 -> lskge3 :155-203
        RandBLAS::sketch_general(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::NoTrans, d, n, m, alpha, S, ro, co, A, lda, beta, B, ldb);
    else        // skge.hh:1031-1052 -> rskge3 :308-356   (here m x d = (m x n)(n x d))
        RandBLAS::sketch_general(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::NoTrans, m, d, n, alpha, A, lda, S, ro, co, beta, B, ldb);
    return 0;
    RL_CATCH
}

// sketch_general with every layout / transposition flag (RandBLAS/RandBLAS/skge.hh:859-905 left, :1031-1076 right); layout: 1 ColMajor, 2 RowMajor
template <typename T>
static int sketch_general_dense_impl(bool left, int layout, int opS, int opA, int64_t S_rows, int64_t S_cols, int family, int major_axis, int64_t d,
                                     int64_t n, int64_t m, T alpha, int64_t ro, int64_t co, const T* A, int64_t lda, T beta, T* B, int64_t ldb,
                                     uint32_t state[6]) {
    RL_TRY
    RandBLAS::DenseDist D(S_rows, S_cols, family == RL_FAMILY_UNIFORM ? RandBLAS::ScalarDist::Uniform : RandBLAS::ScalarDist::Gaussian,
                          major_axis == RL_AXIS_SHORT ? RandBLAS::Axis::Short : RandBLAS::Axis::Long);
    State st = load_state(state);
    RandBLAS::DenseSkOp<T, RNG> S(D, st);
    store_state(S.next_state, state);
    const blas::Layout L = layout == 2 ? blas::Layout::RowMajor : blas::Layout::ColMajor;
    const blas::Op oS = opS ? blas::Op::Trans : blas::Op::NoTrans, oA = opA ? blas::Op::Trans : blas::Op::NoTrans;
    if (left) RandBLAS::sketch_general(L, oS, oA, d, n, m, alpha, S, ro, co, A, lda, beta, B, ldb);
    else      RandBLAS::sketch_general(L, oA, oS, m, d, n, alpha, A, lda, S, ro, co, beta, B, ldb);
    return 0;
    RL_CATCH
}

// the same with a short-axis SparseSkOp (skge.hh:907-960 left, :1078-1131 right)
template <typename T>
static int sketch_general_sparse_impl(bool left, int layout, int opS, int opA, int64_t S_rows, int64_t S_cols, int64_t vec_nnz, int64_t d, int64_t n,
                                      int64_t m, T alpha, int64_t ro, int64_t co, const T* A, int64_t lda, T beta, T* B, int64_t ldb,
                                      uint32_t state[6], int major_axis = RL_AXIS_SHORT) {
    RL_TRY
    RandBLAS::SparseDist DS(S_rows, S_cols, vec_nnz, major_axis == RL_AXIS_SHORT ? RandBLAS::Axis::Short : RandBLAS::Axis::Long);
    State st = load_state(state);
    RandBLAS::SparseSkOp<T, RNG> S(DS, st);
    store_state(S.next_state, state);
    const blas::Layout L = layout == 2 ? blas::Layout::RowMajor : blas::Layout::ColMajor;
    const blas::Op oS = opS ? blas::Op::Trans : blas::Op::NoTrans, oA = opA ? blas::Op::Trans : blas::Op::NoTrans;
    if (left) RandBLAS::sketch_general(L, oS, oA, d, n, m, alpha, S, ro, co, A, lda, beta, B, ldb);
    else      RandBLAS::sketch_general(L, oA, oS, m, d, n, alpha, A, lda, S, ro, co, beta, B, ldb);
    return 0;
    RL_CATCH
}

// CQRRPT (RandLAPACK/drivers/rl_cqrrpt.hh:146-391) with the default subroutines (geqp3) unless qrcp says otherwise
template <typename T>
static int cqrrpt_impl(int64_t m, int64_t n, T* A, int64_t lda, T* R, int64_t ldr, int64_t* J, T d_factor, T eps, int64_t nnz, int qrcp,
                       int64_t* rank, uint32_t state[6], int orthogonalization = 0) {
    RL_TRY
    RandLAPACK::CQRRPT<T, RNG> alg(false, eps);
    alg.nnz = nnz;
    alg.orthogonalization = orthogonalization != 0;
    alg.qrcp = qrcp == 1 ? RandLAPACK::CQRRPTSubroutines::QRCP::bqrrp
             : qrcp == 2 ? RandLAPACK::CQRRPTSubroutines::QRCP::hqrrp : RandLAPACK::CQRRPTSubroutines::QRCP::geqp3;
    State st = load_state(state);
    int rc = alg.call(m, n, A, lda, R, ldr, J, d_factor, st);
    store_state(st, state);
    *rank = alg.rank;
    return rc;
    RL_CATCH
}

// CQRRT (RandLAPACK/drivers/rl_cqrrt.hh:91-297)
template <typename T>
static int cqrrt_impl(int64_t m, int64_t n, T* A, int64_t lda, T* R, int64_t ldr, T d_factor, T eps, int64_t nnz, int orthogonalization,
                      int compute_Q, uint32_t state[6]) {
    RL_TRY
    RandLAPACK::CQRRT<T, RNG> alg(false, eps);
    alg.nnz = nnz;
    alg.orthogonalization = orthogonalization != 0;
    alg.compute_Q = compute_Q != 0;
    State st = load_state(state);
    int rc = alg.call(m, n, A, lda, R, ldr, d_factor, st);
    store_state(st, state);
    return rc;
    RL_CATCH
}

// BQRRP (RandLAPACK/drivers/rl_bqrrp.hh:154-665). qrcp_wide: 0 luqr (default), 1 geqp3; qr_tall: 0 geqrf (default), 1 cholqr, 2 geqrt
template <typename T>
static int bqrrp_impl(int64_t m, int64_t n, T* A, int64_t lda, T d_factor, int64_t b_sz, int qrcp_wide, int qr_tall, T* tau, int64_t* J,
                      int64_t* rank, uint32_t state[6], T tol = 0) {
    RL_TRY
    RandLAPACK::BQRRP<T, RNG> alg(false, b_sz);
    if (tol > 0) alg.tol = tol;
    alg.qrcp_wide = qrcp_wide == 1 ? RandLAPACK::BQRRPSubroutines::QRCPWide::geqp3 : RandLAPACK::BQRRPSubroutines::QRCPWide::luqr;
    alg.qr_tall = qr_tall == 1 ? RandLAPACK::BQRRPSubroutines::QRTall::cholqr
                : qr_tall == 2 ? RandLAPACK::BQRRPSubroutines::QRTall::geqrt : RandLAPACK::BQRRPSubroutines::QRTall::geqrf;
    State st = load_state(state);
    int rc = alg.call(m, n, A, lda, d_factor, tau, J, st);
    store_state(st, state);
    *rank = alg.rank;
    return rc;
    RL_CATCH
}

// hqrrp (RandLAPACK/drivers/rl_hqrrp.hh:811-1196): Householder QR with randomized pivoting; J (1-based) and tau as geqp3 returns them
template <typename T>
static int hqrrp_impl(int64_t m, int64_t n, T* A, int64_t lda, int64_t* J, T* tau, int64_t nb_alg, int64_t pp, int64_t panel_pivoting,
                      int64_t qr_type, uint32_t state[6]) {
    RL_TRY
    State st = load_state(state);
    int rc = (int)RandLAPACK::hqrrp(m, n, A, lda, J, tau, nb_alg, pp, panel_pivoting, qr_type, st, (T**)nullptr);
    store_state(st, state);
    return rc;
    RL_CATCH
}

// SYPS / SYRF / REVD2 (RandLAPACK/comps/rl_syps.hh:21-143, comps/rl_syrf.hh:21-118, drivers/rl_revd2.hh:75-246); the algorithm objects
// of test/drivers/test_revd2.cc:78-101.  uplo: 0 upper, 1 lower.
template <typename T>
static int syps_impl(int uplo, int64_t m, const T* A, int64_t lda, int64_t k, int64_t p, int64_t q, T* skop, uint32_t state[6]) {
    RL_TRY
    RandLAPACK::SYPS<T, RNG> syps(p, q, false, false);
    State st = load_state(state);
    T* sk = skop;
    std::vector<T> work(m * k, 0.0);
    int rc = syps.call(uplo ? blas::Uplo::Lower : blas::Uplo::Upper, m, A, lda, k, st, sk, work.data());
    store_state(st, state);
    return rc;
    RL_CATCH
}
template <typename T>
static int syrf_impl(int uplo, int64_t m, const T* A, int64_t k, int64_t p, int64_t q, int orth, T* Q, uint32_t state[6]) {
    RL_TRY
    RandLAPACK::SYPS<T, RNG> syps(p, q, false, false);
    auto o = make_stab<T>(orth, false);
    RandLAPACK::SYRF<RandLAPACK::SYPS<T, RNG>, RandLAPACK::Stabilization<T>> syrf(syps, *o, false, false);
    State st = load_state(state);
    std::vector<T> Qv;
    int rc = syrf.call(uplo ? blas::Uplo::Lower : blas::Uplo::Upper, m, A, k, Qv, st, nullptr);
    std::memcpy(Q, Qv.data(), sizeof(T) * m * k);
    store_state(st, state);
    return rc;
    RL_CATCH
}
template <typename T>
static int revd2_impl(int uplo, int64_t m, const T* A, int64_t* k, int64_t k_cap, T tol, int64_t p, int64_t q, int orth, int error_est_p,
                      T* V, T* eigvals, uint32_t state[6]) {
    RL_TRY
    RandLAPACK::SYPS<T, RNG> syps(p, q, false, false);
    auto o = make_stab<T>(orth, false);
    using SYRF_t = RandLAPACK::SYRF<RandLAPACK::SYPS<T, RNG>, RandLAPACK::Stabilization<T>>;
    SYRF_t syrf(syps, *o, false, false);
    RandLAPACK::REVD2<SYRF_t> revd2(syrf, error_est_p, false);
    State st = load_state(state);
    std::vector<T> Vv, ev;
    int rc = revd2.call(uplo ? blas::Uplo::Lower : blas::Uplo::Upper, m, A, *k, tol, Vv, ev, st);
    if (*k > k_cap) throw std::runtime_error("revd2: k grew beyond the caller's capacity");
    std::memcpy(V, Vv.data(), sizeof(T) * m * (*k));
    std::memcpy(eigvals, ev.data(), sizeof(T) * (*k));
    store_state(st, state);
    return rc;
    RL_CATCH
}

extern "C" {

const char* rlref_last_error(void) { return g_err; }
const char* rlref_kind(void) { return "reference"; }

int rlref_set_num_threads(int n) {
    scipy_openblas_set_num_threads(n);
#if defined(RandBLAS_HAS_OpenMP)
    omp_set_num_threads(n);
#endif
    return 0;
}
int rlref_get_num_threads(void) { return scipy_openblas_get_num_threads(); }

int rlref_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    RNG rng; RNG::ctr_type c; RNG::key_type k;
    for (int i = 0; i < 4; ++i) c.v[i] = ctr[i];
    for (int i = 0; i < 2; ++i) k.v[i] = key[i];
    auto r = rng(c, k);
    for (int i = 0; i < 4; ++i) out[i] = r.v[i];
    return 0;
}

int rlref_ctr_incr(uint32_t ctr[4], uint64_t step) {
    RNG::ctr_type c;
    for (int i = 0; i < 4; ++i) c.v[i] = ctr[i];
    c.incr(step);
    for (int i = 0; i < 4; ++i) ctr[i] = c.v[i];
    return 0;
}

int rlref_boxmuller(uint32_t u0, uint32_t u1, float out[2]) {
    auto f = r123::boxmuller(u0, u1);
    out[0] = f.x; out[1] = f.y;
    return 0;
}

int rlref_fill_dense_f64(int64_t n_rows, int64_t n_cols, int family, int major_axis, int layout, int64_t sub_rows, int64_t sub_cols,
                         int64_t ro, int64_t co, double* buff, uint32_t state[6]) {
    return fill_dense_impl<double>(n_rows, n_cols, family, major_axis, layout, sub_rows, sub_cols, ro, co, buff, state);
}
int rlref_fill_dense_f32(int64_t n_rows, int64_t n_cols, int family, int major_axis, int layout, int64_t sub_rows, int64_t sub_cols,
                         int64_t ro, int64_t co, float* buff, uint32_t state[6]) {
    return fill_dense_impl<float>(n_rows, n_cols, family, major_axis, layout, sub_rows, sub_cols, ro, co, buff, state);
}

int rlref_stab_f64(int kind, int64_t m, int64_t k, double* A, int cond_check) { return stab_impl<double>(kind, m, k, A, cond_check); }
int rlref_stab_f32(int kind, int64_t m, int64_t k, float* A, int cond_check) { return stab_impl<float>(kind, m, k, A, cond_check); }

int rlref_rs_f64(int64_t m, int64_t n, const double* A, int64_t k, double* Omega, uint32_t state[6], const rl_stack_opts* o) {
    return rs_impl<double>(m, n, A, k, Omega, state, o);
}
int rlref_rs_f32(int64_t m, int64_t n, const float* A, int64_t k, float* Omega, uint32_t state[6], const rl_stack_opts* o) {
    return rs_impl<float>(m, n, A, k, Omega, state, o);
}
int rlref_rf_f64(int64_t m, int64_t n, const double* A, int64_t k, double* Q, uint32_t state[6], const rl_stack_opts* o) {
    return rf_impl<double>(m, n, A, k, Q, state, o);
}
int rlref_rf_f32(int64_t m, int64_t n, const float* A, int64_t k, float* Q, uint32_t state[6], const rl_stack_opts* o) {
    return rf_impl<float>(m, n, A, k, Q, state, o);
}
int rlref_qb_f64(int64_t m, int64_t n, double* A, int64_t* k, int64_t b_sz, double tol, double* Q, double* BT, uint32_t state[6],
                 const rl_stack_opts* o) {
    return qb_impl<double>(m, n, A, k, b_sz, tol, Q, BT, state, o);
}
int rlref_qb_f32(int64_t m, int64_t n, float* A, int64_t* k, int64_t b_sz, float tol, float* Q, float* BT, uint32_t state[6],
                 const rl_stack_opts* o) {
    return qb_impl<float>(m, n, A, k, b_sz, tol, Q, BT, state, o);
}
int rlref_rsvd_f64(int64_t m, int64_t n, double* A, int64_t* k, double tol, double* U, double* S, double* V, uint32_t state[6],
                   const rl_stack_opts* o) {
    return rsvd_impl<double>(m, n, A, k, tol, U, S, V, state, o);
}
int rlref_rsvd_f32(int64_t m, int64_t n, float* A, int64_t* k, float tol, float* U, float* S, float* V, uint32_t state[6],
                   const rl_stack_opts* o) {
    return rsvd_impl<float>(m, n, A, k, tol, U, S, V, state, o);
}

int rlref_mat_gen_f64(int type, int64_t m, int64_t n, int64_t rank, double cond, double exponent, double scaling, double* A,
                      uint32_t state[6]) {
    return mat_gen_impl<double>(type, m, n, rank, cond, exponent, scaling, A, state);
}
int rlref_mat_gen_f32(int type, int64_t m, int64_t n, int64_t rank, float cond, float exponent, float scaling, float* A,
                      uint32_t state[6]) {
    return mat_gen_impl<float>(type, m, n, rank, cond, exponent, scaling, A, state);
}

#define RLREF_TYPED(T, SUF)                                                                                                               \
    int rlref_fill_sparse_##SUF(int64_t n_rows, int64_t n_cols, int64_t vec_nnz, int major_axis, int64_t sub_rows, int64_t sub_cols,       \
                                int64_t ro, int64_t co, int64_t* nnz, T* vals, int64_t* rows, int64_t* cols, uint32_t state[6]) {          \
        return fill_sparse_impl<T>(n_rows, n_cols, vec_nnz, major_axis, sub_rows, sub_cols, ro, co, nnz, vals, rows, cols, state);         \
    }                                                                                                                                     \
    int rlref_sketch_sparse_left_##SUF(int64_t S_rows, int64_t S_cols, int64_t vec_nnz, int64_t d, int64_t n, int64_t m, T alpha,          \
                                       int64_t ro, int64_t co, const T* A, int64_t lda, T beta, T* B, int64_t ldb, uint32_t state[6]) {    \
        return sketch_sparse_left_impl<T>(S_rows, S_cols, vec_nnz, d, n, m, alpha, ro, co, A, lda, beta, B, ldb, state);                   \
    }                                                                                                                                     \
    int rlref_sketch_dense_left_##SUF(int64_t S_rows, int64_t S_cols, int family, int major_axis, int64_t d, int64_t n, int64_t m, T alpha, \
                                      int64_t ro, int64_t co, const T* A, int64_t lda, T beta, T* B, int64_t ldb, uint32_t state[6]) {     \
        return sketch_dense_impl<T>(true, S_rows, S_cols, family, major_axis, d, n, m, alpha, ro, co, A, lda, beta, B, ldb, state);        \
    }                                                                                                                                     \
    int rlref_sketch_dense_right_##SUF(int64_t S_rows, int64_t S_cols, int family, int major_axis, int64_t m, int64_t d, int64_t n,        \
                                       T alpha, const T* A, int64_t lda, int64_t ro, int64_t co, T beta, T* B, int64_t ldb,                \
                                       uint32_t state[6]) {                                                                                \
        return sketch_dense_impl<T>(false, S_rows, S_cols, family, major_axis, d, n, m, alpha, ro, co, A, lda, beta, B, ldb, state);       \
    }                                                                                                                                     \
    int rlref_sketch_general_dense_##SUF(int left, int layout, int opS, int opA, int64_t S_rows, int64_t S_cols, int family, int major_axis,    \
                                         int64_t d, int64_t n, int64_t m, T alpha, int64_t ro, int64_t co, const T* A, int64_t lda, T beta,    \
                                         T* B, int64_t ldb, uint32_t state[6]) {                                                              \
        return sketch_general_dense_impl<T>(left != 0, layout, opS, opA, S_rows, S_cols, family, major_axis, d, n, m, alpha, ro, co, A, lda,   \
                                            beta, B, ldb, state);                                                                             \
    }                                                                                                                                     \
    int rlref_sketch_general_sparse_##SUF(int left, int layout, int opS, int opA, int64_t S_rows, int64_t S_cols, int64_t vec_nnz, int64_t d,  \
                                          int64_t n, int64_t m, T alpha, int64_t ro, int64_t co, const T* A, int64_t lda, T beta, T* B,        \
                                          int64_t ldb, uint32_t state[6]) {                                                                    \
        return sketch_general_sparse_impl<T>(left != 0, layout, opS, opA, S_rows, S_cols, vec_nnz, d, n, m, alpha, ro, co, A, lda, beta, B,    \
                                             ldb, state);                                                                                     \
    }                                                                                                                                     \
    int rlref_sketch_general_sparse_axis_##SUF(int left, int layout, int opS, int opA, int64_t S_rows, int64_t S_cols, int64_t vec_nnz,       \
                                               int major_axis, int64_t d, int64_t n, int64_t m, T alpha, int64_t ro, int64_t co, const T* A,   \
                                               int64_t lda, T beta, T* B, int64_t ldb, uint32_t state[6]) {                                    \
        return sketch_general_sparse_impl<T>(left != 0, layout, opS, opA, S_rows, S_cols, vec_nnz, d, n, m, alpha, ro, co, A, lda, beta, B,    \
                                             ldb, state, major_axis);                                                                         \
    }                                                                                                                                     \
    int rlref_cqrrpt_##SUF(int64_t m, int64_t n, T* A, int64_t lda, T* R, int64_t ldr, int64_t* J, T d_factor, T eps, int64_t nnz,         \
                           int qrcp, int64_t* rank, uint32_t state[6]) {                                                                   \
        return cqrrpt_impl<T>(m, n, A, lda, R, ldr, J, d_factor, eps, nnz, qrcp, rank, state);                                             \
    }                                                                                                                                     \
    int rlref_cqrrpt_orth_##SUF(int64_t m, int64_t n, T* A, int64_t lda, T* R, int64_t ldr, int64_t* J, T d_factor, T eps, int64_t nnz,    \
                                int qrcp, int64_t* rank, uint32_t state[6]) {                                                              \
        return cqrrpt_impl<T>(m, n, A, lda, R, ldr, J, d_factor, eps, nnz, qrcp, rank, state, 1);                                          \
    }                                                                                                                                     \
    int rlref_cqrrt_##SUF(int64_t m, int64_t n, T* A, int64_t lda, T* R, int64_t ldr, T d_factor, T eps, int64_t nnz, int orthogonalization,     \
                          int compute_Q, uint32_t state[6]) {                                                                              \
        return cqrrt_impl<T>(m, n, A, lda, R, ldr, d_factor, eps, nnz, orthogonalization, compute_Q, state);                               \
    }                                                                                                                                     \
    int rlref_syps_##SUF(int uplo, int64_t m, const T* A, int64_t lda, int64_t k, int64_t p, int64_t q, T* skop, uint32_t state[6]) {      \
        return syps_impl<T>(uplo, m, A, lda, k, p, q, skop, state);                                                                        \
    }                                                                                                                                     \
    int rlref_syrf_##SUF(int uplo, int64_t m, const T* A, int64_t k, int64_t p, int64_t q, int orth, T* Q, uint32_t state[6]) {            \
        return syrf_impl<T>(uplo, m, A, k, p, q, orth, Q, state);                                                                          \
    }                                                                                                                                     \
    int rlref_revd2_##SUF(int uplo, int64_t m, const T* A, int64_t* k, int64_t k_cap, T tol, int64_t p, int64_t q, int orth,               \
                          int error_est_p, T* V, T* eigvals, uint32_t state[6]) {                                                          \
        return revd2_impl<T>(uplo, m, A, k, k_cap, tol, p, q, orth, error_est_p, V, eigvals, state);                                       \
    }                                                                                                                                     \
    int rlref_hqrrp_##SUF(int64_t m, int64_t n, T* A, int64_t lda, int64_t* J, T* tau, int64_t nb_alg, int64_t pp, int64_t panel_pivoting, \
                          int64_t qr_type, uint32_t state[6]) {                                                                            \
        return hqrrp_impl<T>(m, n, A, lda, J, tau, nb_alg, pp, panel_pivoting, qr_type, state);                                            \
    }                                                                                                                                     \
    int rlref_bqrrp_##SUF(int64_t m, int64_t n, T* A, int64_t lda, T d_factor, int64_t b_sz, int qrcp_wide, int qr_tall, T* tau,           \
                          int64_t* J, int64_t* rank, uint32_t state[6]) {                                                                  \
        return bqrrp_impl<T>(m, n, A, lda, d_factor, b_sz, qrcp_wide, qr_tall, tau, J, rank, state);                                       \
    }                                                                                                                                     \
    int rlref_bqrrp_tol_##SUF(int64_t m, int64_t n, T* A, int64_t lda, T d_factor, int64_t b_sz, int qrcp_wide, int qr_tall, T tol, T* tau, \
                              int64_t* J, int64_t* rank, uint32_t state[6]) {                                                              \
        return bqrrp_impl<T>(m, n, A, lda, d_factor, b_sz, qrcp_wide, qr_tall, tau, J, rank, state, tol);                                  \
    }
RLREF_TYPED(double, f64)
RLREF_TYPED(float, f32)

} // extern "C"
