// Shim for LAPACK++'s <lapack/fortran.h> — TEST INFRASTRUCTURE ONLY (oracle build).
// Maps the LAPACK_<name> spellings onto the `scipy_`-prefixed LP64 Fortran symbols of the
// OpenBLAS bundled with scipy. dlarfb/dlarf/dgeqrf are called directly (no hidden string
// lengths) by RandLAPACK/drivers/rl_hqrrp.hh:98-161, so those three are declared that way.
#pragma once
#include "config.h"

#define RLL_SYM(name) scipy_##name##_

extern "C" {
// ---- routines the reference calls raw --------------------------------------
void RLL_SYM(dlarfb)(const char*, const char*, const char*, const char*, const lapack_int*, const lapack_int*, const lapack_int*,
                     double*, const lapack_int*, double*, const lapack_int*, double*, const lapack_int*, double*, const lapack_int*);
void RLL_SYM(slarfb)(const char*, const char*, const char*, const char*, const lapack_int*, const lapack_int*, const lapack_int*,
                     float*, const lapack_int*, float*, const lapack_int*, float*, const lapack_int*, float*, const lapack_int*);
void RLL_SYM(dlarf)(const char*, const lapack_int*, const lapack_int*, double*, const lapack_int*, double*, double*, const lapack_int*, double*);
void RLL_SYM(slarf)(const char*, const lapack_int*, const lapack_int*, float*, const lapack_int*, float*, float*, const lapack_int*, float*);
void RLL_SYM(dgeqrf)(const lapack_int*, const lapack_int*, double*, const lapack_int*, double*, double*, const lapack_int*, lapack_int*);
void RLL_SYM(sgeqrf)(const lapack_int*, const lapack_int*, float*, const lapack_int*, float*, float*, const lapack_int*, lapack_int*);
double RLL_SYM(dlamch)(const char*);
float  RLL_SYM(slamch)(const char*);

// ---- routines only the shim's lapack.hh calls (hidden string lengths passed) ----
#define RLL_PROTOS(T, p) \
void RLL_SYM(p##potrf)(const char*, const lapack_int*, T*, const lapack_int*, lapack_int*, size_t); \
void RLL_SYM(p##getrf)(const lapack_int*, const lapack_int*, T*, const lapack_int*, lapack_int*, lapack_int*); \
void RLL_SYM(p##laswp)(const lapack_int*, T*, const lapack_int*, const lapack_int*, const lapack_int*, const lapack_int*, const lapack_int*); \
void RLL_SYM(p##orgqr)(const lapack_int*, const lapack_int*, const lapack_int*, T*, const lapack_int*, const T*, T*, const lapack_int*, lapack_int*); \
void RLL_SYM(p##ormqr)(const char*, const char*, const lapack_int*, const lapack_int*, const lapack_int*, const T*, const lapack_int*, const T*, T*, const lapack_int*, T*, const lapack_int*, lapack_int*, size_t, size_t); \
void RLL_SYM(p##geqp3)(const lapack_int*, const lapack_int*, T*, const lapack_int*, lapack_int*, T*, T*, const lapack_int*, lapack_int*); \
void RLL_SYM(p##geqrt)(const lapack_int*, const lapack_int*, const lapack_int*, T*, const lapack_int*, T*, const lapack_int*, T*, lapack_int*); \
void RLL_SYM(p##gemqrt)(const char*, const char*, const lapack_int*, const lapack_int*, const lapack_int*, const lapack_int*, const T*, const lapack_int*, const T*, const lapack_int*, T*, const lapack_int*, T*, lapack_int*, size_t, size_t); \
void RLL_SYM(p##orhr_col)(const lapack_int*, const lapack_int*, const lapack_int*, T*, const lapack_int*, T*, const lapack_int*, T*, lapack_int*); \
void RLL_SYM(p##larfg)(const lapack_int*, T*, T*, const lapack_int*, T*); \
void RLL_SYM(p##larft)(const char*, const char*, const lapack_int*, const lapack_int*, const T*, const lapack_int*, const T*, T*, const lapack_int*, size_t, size_t); \
void RLL_SYM(p##gesdd)(const char*, const lapack_int*, const lapack_int*, T*, const lapack_int*, T*, T*, const lapack_int*, T*, const lapack_int*, T*, const lapack_int*, lapack_int*, lapack_int*, size_t); \
void RLL_SYM(p##lacpy)(const char*, const lapack_int*, const lapack_int*, const T*, const lapack_int*, T*, const lapack_int*, size_t); \
T    RLL_SYM(p##lange)(const char*, const lapack_int*, const lapack_int*, const T*, const lapack_int*, T*, size_t); \
T    RLL_SYM(p##lansy)(const char*, const char*, const lapack_int*, const T*, const lapack_int*, T*, size_t, size_t); \
void RLL_SYM(p##laset)(const char*, const lapack_int*, const lapack_int*, const T*, const T*, T*, const lapack_int*, size_t); \
void RLL_SYM(p##lapmt)(const lapack_int*, const lapack_int*, const lapack_int*, T*, const lapack_int*, lapack_int*);
RLL_PROTOS(double, d)
RLL_PROTOS(float, s)
}

#define LAPACK_dlarfb RLL_SYM(dlarfb)
#define LAPACK_slarfb RLL_SYM(slarfb)
#define LAPACK_dlarf  RLL_SYM(dlarf)
#define LAPACK_slarf  RLL_SYM(slarf)
#define LAPACK_dgeqrf RLL_SYM(dgeqrf)
#define LAPACK_sgeqrf RLL_SYM(sgeqrf)
#define LAPACK_dlamch RLL_SYM(dlamch)
#define LAPACK_slamch RLL_SYM(slamch)
#define LAPACK_dpotrf RLL_SYM(dpotrf)
#define LAPACK_spotrf RLL_SYM(spotrf)
#define LAPACK_dgetrf RLL_SYM(dgetrf)
#define LAPACK_sgetrf RLL_SYM(sgetrf)
#define LAPACK_dlaswp RLL_SYM(dlaswp)
#define LAPACK_slaswp RLL_SYM(slaswp)
#define LAPACK_dorgqr RLL_SYM(dorgqr)
#define LAPACK_sorgqr RLL_SYM(sorgqr)
#define LAPACK_dormqr RLL_SYM(dormqr)
#define LAPACK_sormqr RLL_SYM(sormqr)
#define LAPACK_dgeqp3 RLL_SYM(dgeqp3)
#define LAPACK_sgeqp3 RLL_SYM(sgeqp3)
#define LAPACK_dgeqrt RLL_SYM(dgeqrt)
#define LAPACK_sgeqrt RLL_SYM(sgeqrt)
#define LAPACK_dgemqrt RLL_SYM(dgemqrt)
#define LAPACK_sgemqrt RLL_SYM(sgemqrt)
#define LAPACK_dorhr_col RLL_SYM(dorhr_col)
#define LAPACK_sorhr_col RLL_SYM(sorhr_col)
#define LAPACK_dlarfg RLL_SYM(dlarfg)
#define LAPACK_slarfg RLL_SYM(slarfg)
#define LAPACK_dlarft RLL_SYM(dlarft)
#define LAPACK_slarft RLL_SYM(slarft)
#define LAPACK_dgesdd RLL_SYM(dgesdd)
#define LAPACK_sgesdd RLL_SYM(sgesdd)
#define LAPACK_dlacpy RLL_SYM(dlacpy)
#define LAPACK_slacpy RLL_SYM(slacpy)
#define LAPACK_dlange RLL_SYM(dlange)
#define LAPACK_slange RLL_SYM(slange)
#define LAPACK_dlansy RLL_SYM(dlansy)
#define LAPACK_slansy RLL_SYM(slansy)
#define LAPACK_dlaset RLL_SYM(dlaset)
#define LAPACK_slaset RLL_SYM(slaset)
#define LAPACK_dlapmt RLL_SYM(dlapmt)
#define LAPACK_slapmt RLL_SYM(slapmt)
