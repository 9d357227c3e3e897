// Shim for BLAS++'s <blas.hh> — TEST INFRASTRUCTURE ONLY (oracle build).
//
// BLAS++ (icl-utk-edu/blaspp) is an un-vendored, un-pinned dependency of the reference
// (install/install.sh:336-339). This header provides exactly the `blas::` surface the
// sketch-and-factor path touches (enumerated by grep over the headers listed in
// oracle/Makefile) and forwards every call to the LP64 Fortran BLAS inside the OpenBLAS
// that ships with scipy (`scipy_`-prefixed symbols). Semantics follow the BLAS++ docs:
// 64-bit dimensions, Layout-aware level-2/3 wrappers, 0-based iamax.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <complex>
#include <stdexcept>
#include <string>
#include <algorithm>
#include <vector>
#include <limits>
#include <cmath>
#include <cassert>
#include <cstring>
#include <cstdio>
#include <iostream>
#include <tuple>
#include <utility>

namespace blas {

enum class Layout : char { ColMajor = 'C', RowMajor = 'R' };
enum class Op     : char { NoTrans = 'N', Trans = 'T', ConjTrans = 'C' };
enum class Uplo   : char { Upper = 'U', Lower = 'L', General = 'G' };
enum class Diag   : char { NonUnit = 'N', Unit = 'U' };
enum class Side   : char { Left = 'L', Right = 'R' };

inline char to_char(Layout v) { return char(v); }
inline char to_char(Op v)     { return char(v); }
inline char to_char(Uplo v)   { return char(v); }
inline char to_char(Diag v)   { return char(v); }
inline char to_char(Side v)   { return char(v); }

class Error : public std::exception {
public:
    Error() {}
    Error(std::string const& msg) : msg_(msg) {}
    const char* what() const noexcept override { return msg_.c_str(); }
private:
    std::string msg_;
};

template <typename T> inline T real(T x) { return x; }
template <typename T> inline T real(std::complex<T> x) { return x.real(); }

using blas_int = int;

inline blas_int to_blas_int(int64_t x, const char* what) {
    if (x > INT32_MAX || x < INT32_MIN) throw Error(std::string("blas shim: dimension overflows LP64 int: ") + what);
    return (blas_int)x;
}
#define RLB_BI(x) ::blas::to_blas_int((x), #x)

} // namespace blas

extern "C" {
#define RLB_F(name) scipy_##name##_
void RLB_F(dgemm)(const char*, const char*, const int*, const int*, const int*, const double*, const double*, const int*, const double*, const int*, const double*, double*, const int*, size_t, size_t);
void RLB_F(sgemm)(const char*, const char*, const int*, const int*, const int*, const float*, const float*, const int*, const float*, const int*, const float*, float*, const int*, size_t, size_t);
void RLB_F(dsyrk)(const char*, const char*, const int*, const int*, const double*, const double*, const int*, const double*, double*, const int*, size_t, size_t);
void RLB_F(ssyrk)(const char*, const char*, const int*, const int*, const float*, const float*, const int*, const float*, float*, const int*, size_t, size_t);
void RLB_F(dsymm)(const char*, const char*, const int*, const int*, const double*, const double*, const int*, const double*, const int*, const double*, double*, const int*, size_t, size_t);
void RLB_F(ssymm)(const char*, const char*, const int*, const int*, const float*, const float*, const int*, const float*, const int*, const float*, float*, const int*, size_t, size_t);
void RLB_F(dtrsm)(const char*, const char*, const char*, const char*, const int*, const int*, const double*, const double*, const int*, double*, const int*, size_t, size_t, size_t, size_t);
void RLB_F(strsm)(const char*, const char*, const char*, const char*, const int*, const int*, const float*, const float*, const int*, float*, const int*, size_t, size_t, size_t, size_t);
void RLB_F(dtrmm)(const char*, const char*, const char*, const char*, const int*, const int*, const double*, const double*, const int*, double*, const int*, size_t, size_t, size_t, size_t);
void RLB_F(strmm)(const char*, const char*, const char*, const char*, const int*, const int*, const float*, const float*, const int*, float*, const int*, size_t, size_t, size_t, size_t);
void RLB_F(dcopy)(const int*, const double*, const int*, double*, const int*);
void RLB_F(scopy)(const int*, const float*, const int*, float*, const int*);
void RLB_F(dscal)(const int*, const double*, double*, const int*);
void RLB_F(sscal)(const int*, const float*, float*, const int*);
void RLB_F(daxpy)(const int*, const double*, const double*, const int*, double*, const int*);
void RLB_F(saxpy)(const int*, const float*, const float*, const int*, float*, const int*);
void RLB_F(dswap)(const int*, double*, const int*, double*, const int*);
void RLB_F(sswap)(const int*, float*, const int*, float*, const int*);
double RLB_F(dnrm2)(const int*, const double*, const int*);
float  RLB_F(snrm2)(const int*, const float*, const int*);
double RLB_F(ddot)(const int*, const double*, const int*, const double*, const int*);
float  RLB_F(sdot)(const int*, const float*, const int*, const float*, const int*);
int RLB_F(idamax)(const int*, const double*, const int*);
int RLB_F(isamax)(const int*, const float*, const int*);
void RLB_F(dger)(const int*, const int*, const double*, const double*, const int*, const double*, const int*, double*, const int*);
void RLB_F(sger)(const int*, const int*, const float*, const float*, const int*, const float*, const int*, float*, const int*);
void RLB_F(dgemv)(const char*, const int*, const int*, const double*, const double*, const int*, const double*, const int*, const double*, double*, const int*, size_t);
void RLB_F(sgemv)(const char*, const int*, const int*, const float*, const float*, const int*, const float*, const int*, const float*, float*, const int*, size_t);
void scipy_openblas_set_num_threads(int);
int  scipy_openblas_get_num_threads(void);
}

namespace blas {

// ---- level 3 ---------------------------------------------------------------
#define RLB_GEMM(T, f)                                                                                   \
inline void gemm(Layout layout, Op transA, Op transB, int64_t m, int64_t n, int64_t k, T alpha,        \
                 const T* A, int64_t lda, const T* B, int64_t ldb, T beta, T* C, int64_t ldc) {          \
    char ta = to_char(transA), tb = to_char(transB);                                                     \
    int m_ = RLB_BI(m), n_ = RLB_BI(n), k_ = RLB_BI(k), lda_ = RLB_BI(lda), ldb_ = RLB_BI(ldb), ldc_ = RLB_BI(ldc); \
    if (layout == Layout::RowMajor) /* C^T = op(B)^T op(A)^T in column-major */                         \
        RLB_F(f)(&tb, &ta, &n_, &m_, &k_, &alpha, B, &ldb_, A, &lda_, &beta, C, &ldc_, 1, 1);            \
    else                                                                                                 \
        RLB_F(f)(&ta, &tb, &m_, &n_, &k_, &alpha, A, &lda_, B, &ldb_, &beta, C, &ldc_, 1, 1);            \
}
RLB_GEMM(double, dgemm)
RLB_GEMM(float, sgemm)

#define RLB_SYRK(T, f)                                                                                   \
inline void syrk(Layout layout, Uplo uplo, Op trans, int64_t n, int64_t k, T alpha, const T* A,         \
                 int64_t lda, T beta, T* C, int64_t ldc) {                                               \
    if (layout == Layout::RowMajor) {                                                                    \
        uplo  = (uplo == Uplo::Lower ? Uplo::Upper : Uplo::Lower);                                       \
        trans = (trans == Op::NoTrans ? Op::Trans : Op::NoTrans);                                        \
    }                                                                                                    \
    char ul = to_char(uplo), tr = to_char(trans);                                                        \
    int n_ = RLB_BI(n), k_ = RLB_BI(k), lda_ = RLB_BI(lda), ldc_ = RLB_BI(ldc);                          \
    RLB_F(f)(&ul, &tr, &n_, &k_, &alpha, A, &lda_, &beta, C, &ldc_, 1, 1);                               \
}
RLB_SYRK(double, dsyrk)
RLB_SYRK(float, ssyrk)

#define RLB_SYMM(T, f)                                                                                   \
inline void symm(Layout layout, Side side, Uplo uplo, int64_t m, int64_t n, T alpha, const T* A,        \
                 int64_t lda, const T* B, int64_t ldb, T beta, T* C, int64_t ldc) {                      \
    int m_ = RLB_BI(m), n_ = RLB_BI(n), lda_ = RLB_BI(lda), ldb_ = RLB_BI(ldb), ldc_ = RLB_BI(ldc);      \
    if (layout == Layout::RowMajor) {                                                                    \
        side = (side == Side::Left ? Side::Right : Side::Left);                                          \
        uplo = (uplo == Uplo::Lower ? Uplo::Upper : Uplo::Lower);                                        \
        std::swap(m_, n_);                                                                               \
    }                                                                                                    \
    char sd = to_char(side), ul = to_char(uplo);                                                         \
    RLB_F(f)(&sd, &ul, &m_, &n_, &alpha, A, &lda_, B, &ldb_, &beta, C, &ldc_, 1, 1);                     \
}
RLB_SYMM(double, dsymm)
RLB_SYMM(float, ssymm)

#define RLB_TRXM(name, T, f)                                                                             \
inline void name(Layout layout, Side side, Uplo uplo, Op trans, Diag diag, int64_t m, int64_t n,        \
                 T alpha, const T* A, int64_t lda, T* B, int64_t ldb) {                                  \
    int m_ = RLB_BI(m), n_ = RLB_BI(n), lda_ = RLB_BI(lda), ldb_ = RLB_BI(ldb);                          \
    if (layout == Layout::RowMajor) {                                                                    \
        side = (side == Side::Left ? Side::Right : Side::Left);                                          \
        uplo = (uplo == Uplo::Lower ? Uplo::Upper : Uplo::Lower);                                        \
        std::swap(m_, n_);                                                                               \
    }                                                                                                    \
    char sd = to_char(side), ul = to_char(uplo), tr = to_char(trans), dg = to_char(diag);                \
    RLB_F(f)(&sd, &ul, &tr, &dg, &m_, &n_, &alpha, A, &lda_, B, &ldb_, 1, 1, 1, 1);                      \
}
RLB_TRXM(trsm, double, dtrsm)
RLB_TRXM(trsm, float, strsm)
RLB_TRXM(trmm, double, dtrmm)
RLB_TRXM(trmm, float, strmm)

// ---- level 2 ---------------------------------------------------------------
#define RLB_GER(T, f)                                                                                    \
inline void ger(Layout layout, int64_t m, int64_t n, T alpha, const T* x, int64_t incx, const T* y,     \
                int64_t incy, T* A, int64_t lda) {                                                       \
    int m_ = RLB_BI(m), n_ = RLB_BI(n), incx_ = RLB_BI(incx), incy_ = RLB_BI(incy), lda_ = RLB_BI(lda);  \
    if (layout == Layout::RowMajor) RLB_F(f)(&n_, &m_, &alpha, y, &incy_, x, &incx_, A, &lda_);          \
    else                            RLB_F(f)(&m_, &n_, &alpha, x, &incx_, y, &incy_, A, &lda_);          \
}
RLB_GER(double, dger)
RLB_GER(float, sger)

#define RLB_GEMV(T, f)                                                                                   \
inline void gemv(Layout layout, Op trans, int64_t m, int64_t n, T alpha, const T* A, int64_t lda,       \
                 const T* x, int64_t incx, T beta, T* y, int64_t incy) {                                 \
    int m_ = RLB_BI(m), n_ = RLB_BI(n), incx_ = RLB_BI(incx), incy_ = RLB_BI(incy), lda_ = RLB_BI(lda);  \
    if (layout == Layout::RowMajor) {                                                                    \
        trans = (trans == Op::NoTrans ? Op::Trans : Op::NoTrans);                                        \
        std::swap(m_, n_);                                                                               \
    }                                                                                                    \
    char tr = to_char(trans);                                                                            \
    RLB_F(f)(&tr, &m_, &n_, &alpha, A, &lda_, x, &incx_, &beta, y, &incy_, 1);                           \
}
RLB_GEMV(double, dgemv)
RLB_GEMV(float, sgemv)

// ---- level 1 ---------------------------------------------------------------
#define RLB_L1(T, p)                                                                                     \
inline void copy(int64_t n, const T* x, int64_t incx, T* y, int64_t incy) {                              \
    int n_ = RLB_BI(n), ix = RLB_BI(incx), iy = RLB_BI(incy); RLB_F(p##copy)(&n_, x, &ix, y, &iy); }     \
inline void scal(int64_t n, T alpha, T* x, int64_t incx) {                                               \
    int n_ = RLB_BI(n), ix = RLB_BI(incx); RLB_F(p##scal)(&n_, &alpha, x, &ix); }                        \
inline void axpy(int64_t n, T alpha, const T* x, int64_t incx, T* y, int64_t incy) {                     \
    int n_ = RLB_BI(n), ix = RLB_BI(incx), iy = RLB_BI(incy); RLB_F(p##axpy)(&n_, &alpha, x, &ix, y, &iy); } \
inline void swap(int64_t n, T* x, int64_t incx, T* y, int64_t incy) {                                    \
    int n_ = RLB_BI(n), ix = RLB_BI(incx), iy = RLB_BI(incy); RLB_F(p##swap)(&n_, x, &ix, y, &iy); }     \
inline T nrm2(int64_t n, const T* x, int64_t incx) {                                                     \
    int n_ = RLB_BI(n), ix = RLB_BI(incx); return RLB_F(p##nrm2)(&n_, x, &ix); }                         \
inline T dot(int64_t n, const T* x, int64_t incx, const T* y, int64_t incy) {                            \
    int n_ = RLB_BI(n), ix = RLB_BI(incx), iy = RLB_BI(incy); return RLB_F(p##dot)(&n_, x, &ix, y, &iy); } \
inline int64_t iamax(int64_t n, const T* x, int64_t incx) {                                              \
    int n_ = RLB_BI(n), ix = RLB_BI(incx); return (int64_t)RLB_F(i##p##amax)(&n_, x, &ix) - 1; }
RLB_L1(double, d)
RLB_L1(float, s)

// BLAS++ also ships generic template fallbacks; the path uses copy/swap on int64_t index vectors
// (rl_bqrrp.hh:386,403; rl_cqrrpt.hh:287).
template <typename TX, typename TY>
inline void copy(int64_t n, const TX* x, int64_t incx, TY* y, int64_t incy) {
    int64_t ix = (incx >= 0 ? 0 : (1 - n) * incx), iy = (incy >= 0 ? 0 : (1 - n) * incy);
    for (int64_t i = 0; i < n; ++i, ix += incx, iy += incy) y[iy] = (TY)x[ix];
}
template <typename TX, typename TY>
inline void swap(int64_t n, TX* x, int64_t incx, TY* y, int64_t incy) {
    int64_t ix = (incx >= 0 ? 0 : (1 - n) * incx), iy = (incy >= 0 ? 0 : (1 - n) * incy);
    for (int64_t i = 0; i < n; ++i, ix += incx, iy += incy) std::swap(x[ix], y[iy]);
}

} // namespace blas
