// Shim for LAPACK++'s <lapack.hh> — TEST INFRASTRUCTURE ONLY (oracle build).
//
// LAPACK++ (icl-utk-edu/lapackpp) is an un-vendored, un-pinned dependency of the reference
// (install/install.sh:336-339). This header provides exactly the `lapack::` surface the
// sketch-and-factor path touches and forwards to the LP64 Fortran LAPACK inside scipy's
// bundled OpenBLAS (`scipy_`-prefixed symbols). Semantics follow the LAPACK++ docs: 64-bit
// dimensions and pivot vectors, internal workspace queries, `info` returned as int64_t.
#pragma once
#include "blas.hh"
#include "lapack/fortran.h"
#include <vector>

namespace lapack {

using blas::Layout;
using blas::Op;
using blas::Uplo;
using blas::Diag;
using blas::Side;
using blas::Error;

enum class Job : char { NoVec = 'N', Vec = 'V', UpdateVec = 'U', AllVec = 'A', SomeVec = 'S', OverwriteVec = 'O' };
enum class Norm : char { One = '1', Two = '2', Inf = 'I', Fro = 'F', Max = 'M' };
enum class MatrixType : char { General = 'G', Lower = 'L', Upper = 'U', Hessenberg = 'H', LowerBand = 'B', UpperBand = 'Q', Band = 'Z' };
enum class Direction : char { Forward = 'F', Backward = 'B' };
enum class StoreV : char { Columnwise = 'C', Rowwise = 'R' };

inline char to_char(Job v)        { return char(v); }
inline char to_char(Norm v)       { return char(v); }
inline char to_char(MatrixType v) { return char(v); }
inline char to_char(Direction v)  { return char(v); }
inline char to_char(StoreV v)     { return char(v); }
using blas::to_char;

#define RLL_I(x) ::blas::to_blas_int((x), #x)

// ---- Cholesky / LU ----------------------------------------------------------
#define RLL_POTRF(T, p)                                                                         \
inline int64_t potrf(Uplo uplo, int64_t n, T* A, int64_t lda) {                                 \
    char ul = to_char(uplo); int n_ = RLL_I(n), lda_ = RLL_I(lda), info = 0;                    \
    LAPACK_##p##potrf(&ul, &n_, A, &lda_, &info, 1); return info; }
RLL_POTRF(double, d) RLL_POTRF(float, s)

#define RLL_GETRF(T, p)                                                                         \
inline int64_t getrf(int64_t m, int64_t n, T* A, int64_t lda, int64_t* ipiv) {                  \
    int m_ = RLL_I(m), n_ = RLL_I(n), lda_ = RLL_I(lda), info = 0;                              \
    std::vector<int> ip(std::max<int64_t>(1, std::min(m, n)));                                  \
    LAPACK_##p##getrf(&m_, &n_, A, &lda_, ip.data(), &info);                                    \
    for (int64_t i = 0; i < std::min(m, n); ++i) ipiv[i] = ip[i];                               \
    return info; }
RLL_GETRF(double, d) RLL_GETRF(float, s)

#define RLL_LASWP(T, p)                                                                         \
inline void laswp(int64_t n, T* A, int64_t lda, int64_t k1, int64_t k2, const int64_t* ipiv, int64_t incx) { \
    int n_ = RLL_I(n), lda_ = RLL_I(lda), k1_ = RLL_I(k1), k2_ = RLL_I(k2), incx_ = RLL_I(incx); \
    int64_t len = k1 + (k2 - k1) * std::abs(incx);                                              \
    std::vector<int> ip(std::max<int64_t>(1, len));                                             \
    for (int64_t i = 0; i < len; ++i) ip[i] = (int)ipiv[i];                                     \
    LAPACK_##p##laswp(&n_, A, &lda_, &k1_, &k2_, ip.data(), &incx_); }
RLL_LASWP(double, d) RLL_LASWP(float, s)

// ---- QR family --------------------------------------------------------------
#define RLL_GEQRF(T, p)                                                                         \
inline int64_t geqrf(int64_t m, int64_t n, T* A, int64_t lda, T* tau) {                         \
    int m_ = RLL_I(m), n_ = RLL_I(n), lda_ = RLL_I(lda), info = 0, lwork = -1; T q;             \
    LAPACK_##p##geqrf(&m_, &n_, A, &lda_, tau, &q, &lwork, &info);                              \
    if (info) return info;                                                                      \
    lwork = std::max(1, (int)q); std::vector<T> w(lwork);                                       \
    LAPACK_##p##geqrf(&m_, &n_, A, &lda_, tau, w.data(), &lwork, &info); return info; }
RLL_GEQRF(double, d) RLL_GEQRF(float, s)

#define RLL_ORGQR(T, p)                                                                         \
inline int64_t orgqr(int64_t m, int64_t n, int64_t k, T* A, int64_t lda, const T* tau) {        \
    int m_ = RLL_I(m), n_ = RLL_I(n), k_ = RLL_I(k), lda_ = RLL_I(lda), info = 0, lwork = -1; T q; \
    LAPACK_##p##orgqr(&m_, &n_, &k_, A, &lda_, tau, &q, &lwork, &info);                         \
    if (info) return info;                                                                      \
    lwork = std::max(1, (int)q); std::vector<T> w(lwork);                                       \
    LAPACK_##p##orgqr(&m_, &n_, &k_, A, &lda_, tau, w.data(), &lwork, &info); return info; }    \
inline int64_t ungqr(int64_t m, int64_t n, int64_t k, T* A, int64_t lda, const T* tau) {        \
    return orgqr(m, n, k, A, lda, tau); }
RLL_ORGQR(double, d) RLL_ORGQR(float, s)

#define RLL_ORMQR(T, p)                                                                         \
inline int64_t ormqr(Side side, Op trans, int64_t m, int64_t n, int64_t k, const T* A, int64_t lda, \
                     const T* tau, T* C, int64_t ldc) {                                         \
    char sd = to_char(side), tr = to_char(trans);                                               \
    int m_ = RLL_I(m), n_ = RLL_I(n), k_ = RLL_I(k), lda_ = RLL_I(lda), ldc_ = RLL_I(ldc), info = 0, lwork = -1; T q; \
    LAPACK_##p##ormqr(&sd, &tr, &m_, &n_, &k_, A, &lda_, tau, C, &ldc_, &q, &lwork, &info, 1, 1); \
    if (info) return info;                                                                      \
    lwork = std::max(1, (int)q); std::vector<T> w(lwork);                                       \
    LAPACK_##p##ormqr(&sd, &tr, &m_, &n_, &k_, A, &lda_, tau, C, &ldc_, w.data(), &lwork, &info, 1, 1); \
    return info; }                                                                              \
inline int64_t unmqr(Side side, Op trans, int64_t m, int64_t n, int64_t k, const T* A, int64_t lda, \
                     const T* tau, T* C, int64_t ldc) { return ormqr(side, trans, m, n, k, A, lda, tau, C, ldc); }
RLL_ORMQR(double, d) RLL_ORMQR(float, s)

#define RLL_GEQP3(T, p)                                                                         \
inline int64_t geqp3(int64_t m, int64_t n, T* A, int64_t lda, int64_t* jpvt, T* tau) {          \
    int m_ = RLL_I(m), n_ = RLL_I(n), lda_ = RLL_I(lda), info = 0, lwork = -1; T q;             \
    std::vector<int> jp(std::max<int64_t>(1, n));                                               \
    for (int64_t i = 0; i < n; ++i) jp[i] = (int)jpvt[i];                                       \
    LAPACK_##p##geqp3(&m_, &n_, A, &lda_, jp.data(), tau, &q, &lwork, &info);                   \
    if (info) return info;                                                                      \
    lwork = std::max(1, (int)q); std::vector<T> w(lwork);                                       \
    LAPACK_##p##geqp3(&m_, &n_, A, &lda_, jp.data(), tau, w.data(), &lwork, &info);             \
    for (int64_t i = 0; i < n; ++i) jpvt[i] = jp[i];                                            \
    return info; }
RLL_GEQP3(double, d) RLL_GEQP3(float, s)

#define RLL_GEQRT(T, p)                                                                         \
inline int64_t geqrt(int64_t m, int64_t n, int64_t nb, T* A, int64_t lda, T* Tm, int64_t ldt) { \
    int m_ = RLL_I(m), n_ = RLL_I(n), nb_ = RLL_I(nb), lda_ = RLL_I(lda), ldt_ = RLL_I(ldt), info = 0; \
    std::vector<T> w(std::max<int64_t>(1, nb * n));                                             \
    LAPACK_##p##geqrt(&m_, &n_, &nb_, A, &lda_, Tm, &ldt_, w.data(), &info); return info; }     \
inline int64_t gemqrt(Side side, Op trans, int64_t m, int64_t n, int64_t k, int64_t nb, const T* V, \
                      int64_t ldv, const T* Tm, int64_t ldt, T* C, int64_t ldc) {               \
    char sd = to_char(side), tr = to_char(trans);                                               \
    int m_ = RLL_I(m), n_ = RLL_I(n), k_ = RLL_I(k), nb_ = RLL_I(nb), ldv_ = RLL_I(ldv), ldt_ = RLL_I(ldt), ldc_ = RLL_I(ldc), info = 0; \
    std::vector<T> w(std::max<int64_t>(1, (side == Side::Left ? n : m) * nb));                  \
    LAPACK_##p##gemqrt(&sd, &tr, &m_, &n_, &k_, &nb_, V, &ldv_, Tm, &ldt_, C, &ldc_, w.data(), &info, 1, 1); \
    return info; }                                                                              \
inline int64_t orhr_col(int64_t m, int64_t n, int64_t nb, T* A, int64_t lda, T* Tm, int64_t ldt, T* D) { \
    int m_ = RLL_I(m), n_ = RLL_I(n), nb_ = RLL_I(nb), lda_ = RLL_I(lda), ldt_ = RLL_I(ldt), info = 0; \
    LAPACK_##p##orhr_col(&m_, &n_, &nb_, A, &lda_, Tm, &ldt_, D, &info); return info; }         \
inline int64_t unhr_col(int64_t m, int64_t n, int64_t nb, T* A, int64_t lda, T* Tm, int64_t ldt, T* D) { \
    return orhr_col(m, n, nb, A, lda, Tm, ldt, D); }
RLL_GEQRT(double, d) RLL_GEQRT(float, s)

#define RLL_LARF(T, p)                                                                          \
inline void larfg(int64_t n, T* alpha, T* x, int64_t incx, T* tau) {                            \
    int n_ = RLL_I(n), incx_ = RLL_I(incx); LAPACK_##p##larfg(&n_, alpha, x, &incx_, tau); }    \
inline void larft(Direction direction, StoreV storev, int64_t n, int64_t k, const T* V, int64_t ldv, \
                  const T* tau, T* Tm, int64_t ldt) {                                           \
    char d = to_char(direction), s = to_char(storev);                                           \
    int n_ = RLL_I(n), k_ = RLL_I(k), ldv_ = RLL_I(ldv), ldt_ = RLL_I(ldt);                     \
    LAPACK_##p##larft(&d, &s, &n_, &k_, V, &ldv_, tau, Tm, &ldt_, 1, 1); }                      \
inline void larfb(Side side, Op trans, Direction direction, StoreV storev, int64_t m, int64_t n, int64_t k, \
                  const T* V, int64_t ldv, const T* Tm, int64_t ldt, T* C, int64_t ldc) {       \
    char sd = to_char(side), tr = to_char(trans), d = to_char(direction), s = to_char(storev);  \
    int m_ = RLL_I(m), n_ = RLL_I(n), k_ = RLL_I(k), ldv_ = RLL_I(ldv), ldt_ = RLL_I(ldt), ldc_ = RLL_I(ldc); \
    int ldw = (side == Side::Left ? n_ : m_); if (ldw < 1) ldw = 1;                             \
    std::vector<T> w((size_t)ldw * std::max(1, k_));                                            \
    LAPACK_##p##larfb(&sd, &tr, &d, &s, &m_, &n_, &k_, (T*)V, &ldv_, (T*)Tm, &ldt_, C, &ldc_, w.data(), &ldw); }
RLL_LARF(double, d) RLL_LARF(float, s)

// ---- SVD --------------------------------------------------------------------
#define RLL_GESDD(T, p)                                                                         \
inline int64_t gesdd(Job jobz, int64_t m, int64_t n, T* A, int64_t lda, T* S, T* U, int64_t ldu, \
                     T* VT, int64_t ldvt) {                                                     \
    char jz = to_char(jobz);                                                                    \
    int m_ = RLL_I(m), n_ = RLL_I(n), lda_ = RLL_I(lda), ldu_ = RLL_I(ldu), ldvt_ = RLL_I(ldvt), info = 0, lwork = -1; T q; \
    std::vector<int> iwork(8 * std::max<int64_t>(1, std::min(m, n)));                           \
    LAPACK_##p##gesdd(&jz, &m_, &n_, A, &lda_, S, U, &ldu_, VT, &ldvt_, &q, &lwork, iwork.data(), &info, 1); \
    if (info) return info;                                                                      \
    lwork = std::max(1, (int)q); std::vector<T> w(lwork);                                       \
    LAPACK_##p##gesdd(&jz, &m_, &n_, A, &lda_, S, U, &ldu_, VT, &ldvt_, w.data(), &lwork, iwork.data(), &info, 1); \
    return info; }
RLL_GESDD(double, d) RLL_GESDD(float, s)

// ---- auxiliaries ------------------------------------------------------------
#define RLL_AUX(T, p)                                                                           \
inline void lacpy(MatrixType mt, int64_t m, int64_t n, const T* A, int64_t lda, T* B, int64_t ldb) { \
    char u = to_char(mt); int m_ = RLL_I(m), n_ = RLL_I(n), lda_ = RLL_I(lda), ldb_ = RLL_I(ldb); \
    LAPACK_##p##lacpy(&u, &m_, &n_, A, &lda_, B, &ldb_, 1); }                                   \
inline T lange(Norm norm, int64_t m, int64_t n, const T* A, int64_t lda) {                      \
    char nm = to_char(norm); int m_ = RLL_I(m), n_ = RLL_I(n), lda_ = RLL_I(lda);               \
    std::vector<T> w(norm == Norm::Inf ? std::max<int64_t>(1, m) : 1);                          \
    return LAPACK_##p##lange(&nm, &m_, &n_, A, &lda_, w.data(), 1); }                           \
inline T lansy(Norm norm, Uplo uplo, int64_t n, const T* A, int64_t lda) {                      \
    char nm = to_char(norm), ul = to_char(uplo); int n_ = RLL_I(n), lda_ = RLL_I(lda);          \
    std::vector<T> w(std::max<int64_t>(1, n));                                                  \
    return LAPACK_##p##lansy(&nm, &ul, &n_, A, &lda_, w.data(), 1, 1); }                        \
inline void laset(MatrixType mt, int64_t m, int64_t n, T offdiag, T diag, T* A, int64_t lda) {  \
    char u = to_char(mt); int m_ = RLL_I(m), n_ = RLL_I(n), lda_ = RLL_I(lda);                  \
    LAPACK_##p##laset(&u, &m_, &n_, &offdiag, &diag, A, &lda_, 1); }                            \
inline void lapmt(bool forwrd, int64_t m, int64_t n, T* X, int64_t ldx, int64_t* K) {           \
    int fw = forwrd ? 1 : 0, m_ = RLL_I(m), n_ = RLL_I(n), ldx_ = RLL_I(ldx);                   \
    std::vector<int> k(std::max<int64_t>(1, n));                                                \
    for (int64_t i = 0; i < n; ++i) k[i] = (int)K[i];                                           \
    LAPACK_##p##lapmt(&fw, &m_, &n_, X, &ldx_, k.data());                                       \
    for (int64_t i = 0; i < n; ++i) K[i] = k[i]; }
RLL_AUX(double, d) RLL_AUX(float, s)

} // namespace lapack
