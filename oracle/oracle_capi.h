/* Shared declarations for the two CPU checkers (TEST INFRASTRUCTURE ONLY):
 *   rlref_*  — oracle/_ref/librl_ref.so, the unmodified reference headers (oracle/ref_capi.cc)
 *   rlo_*    — oracle/librl_oracle.so, our independent restatement (oracle/rl_oracle.c)
 * The enums and rl_stack_opts are value-compatible with include/rlb200.h so parity tests can
 * drive the product and the checkers with the same arguments. */
#ifndef RL_ORACLE_CAPI_H
#define RL_ORACLE_CAPI_H
#include <stdint.h>

enum { RL_STAB_PLUL = 0, RL_STAB_CHOLQRQ = 1, RL_STAB_HQRQ = 2 };
enum { RL_FAMILY_GAUSSIAN = 0, RL_FAMILY_UNIFORM = 1 };
enum { RL_AXIS_LONG = 0, RL_AXIS_SHORT = 1 };
enum { RL_LAYOUT_NATURAL = 0, RL_LAYOUT_COLMAJOR = 1, RL_LAYOUT_ROWMAJOR = 2 };
enum { RL_ERR_EXCEPTION = -1000 };

typedef struct rl_stack_opts {
    int64_t passes_over_data;  /* RS p   (rl_rs.hh:45) */
    int64_t passes_per_stab;   /* RS q   (rl_rs.hh:46) */
    int64_t block_sz;          /* RSVD block_sz (rl_rsvd.hh:44) */
    int32_t stab;              /* RS stabiliser:   RL_STAB_* */
    int32_t orth_rf;           /* RF orthogonaliser */
    int32_t orth_qb;           /* QB re-orthogonaliser */
    int32_t cond_check;        /* bool */
    int32_t orth_check;        /* bool (QB) */
    int32_t reserved;
} rl_stack_opts;

#endif
