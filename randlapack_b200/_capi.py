"""ctypes binding of librlb200.so (the C-ABI declared in include/rlb200.h).

The library is the product; this module only loads it and declares argument types.  There is no
Python/CPU fallback: if the shared library is missing the import raises, and every compute call
returns RLB200_ERR_CUDA (raised as RuntimeError) when no sm_100 device is usable.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librlb200.so")

OK, ERR_ARG, ERR_CUDA, ERR_ALLOC, ERR_COLLECTIVE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
STAB_PLUL, STAB_CHOLQRQ, STAB_HQRQ = 0, 1, 2
FAMILY_GAUSSIAN, FAMILY_UNIFORM = 0, 1
AXIS_LONG, AXIS_SHORT = 0, 1
LAYOUT_NATURAL, LAYOUT_COLMAJOR, LAYOUT_ROWMAJOR = 0, 1, 2
QRCP_LUQR, QRCP_GEQP3 = 0, 1
QRTALL_GEQRF, QRTALL_CHOLQR, QRTALL_GEQRT = 0, 1, 2
TIMER_GEMM_NN, TIMER_GEMM_TN, TIMER_RIGHTMUL, TIMER_SMALL, TIMER_FILL, TIMER_SKETCH, TIMER_FACTOR = 0, 1, 2, 3, 4, 5, 6

c_i64, c_i32, c_u32, c_int, c_vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p
P_u32, P_i64, P_int = ctypes.POINTER(c_u32), ctypes.POINTER(c_i64), ctypes.POINTER(c_int)


class StackOpts(ctypes.Structure):
    """rlb200_stack_opts (include/rlb200.h) — the canonical stack of test/drivers/test_rsvd.cc:68-93."""
    _fields_ = [("passes_over_data", c_i64), ("passes_per_stab", c_i64), ("block_sz", c_i64), ("stab", c_i32),
                ("orth_rf", c_i32), ("orth_qb", c_i32), ("cond_check", c_i32), ("orth_check", c_i32), ("reserved", c_i32)]


class Revd2Opts(ctypes.Structure):
    """rlb200_revd2_opts (include/rlb200.h) — SYPS(p, q) / SYRF(syps, orth) / REVD2(syrf, error_est_p), test/drivers/test_revd2.cc:78-101."""
    _fields_ = [("syps_passes", c_i64), ("syps_passes_per_stab", c_i64), ("orth", c_i32), ("error_est_p", c_i32)]


UPLO_UPPER, UPLO_LOWER = 0, 1
CQRRPT_QRCP_GEQP3, CQRRPT_QRCP_BQRRP, CQRRPT_QRCP_HQRRP = 0, 1, 2

ALLREDUCE_FN = ctypes.CFUNCTYPE(c_int, c_vp, c_vp, c_i64, c_i32, c_vp)

# name -> (restype, argtypes); mirrors include/rlb200.h one to one (tests/test_abi.py checks the header against this table)
_F = lambda ft: {  # noqa: E731  typed entry points, ft = ctypes float type
    "fill_dense": (c_int, [c_vp, c_i64, c_i64, c_int, c_int, c_int, c_i64, c_i64, c_i64, c_i64, c_vp, P_u32]),
    "gemm": (c_int, [c_vp, c_int, c_int, c_i64, c_i64, c_i64, ft, c_vp, c_i64, c_vp, c_i64, ft, c_vp, c_i64]),
    "stab": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_int, P_int]),
    "rs": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp, P_u32, ctypes.POINTER(StackOpts)]),
    "rf": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, P_u32, ctypes.POINTER(StackOpts)]),
    "qb": (c_int, [c_vp, c_i64, c_i64, c_vp, P_i64, c_i64, ft, c_vp, c_vp, c_vp, P_u32, ctypes.POINTER(StackOpts)]),
    "rsvd": (c_int, [c_vp, c_i64, c_i64, c_vp, P_i64, ft, c_vp, c_vp, c_vp, c_vp, P_u32, ctypes.POINTER(StackOpts), P_int]),
    "svd_tall": (c_int, [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "cqrrpt": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, ft, ft, c_i64, P_i64, P_u32]),
    "bqrrp": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, ft, c_i64, c_int, c_int, c_vp, c_vp, P_i64, P_u32]),
    "qr_small": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp]),
    "col_swap": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp]),
    "fill_sparse": (c_int, [c_vp, c_i64, c_i64, c_i64, c_int, c_i64, c_i64, c_i64, c_i64, P_i64, c_vp, c_vp, c_vp, P_u32]),
    "sketch_sparse_left": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, ft, c_i64, c_i64, c_vp, c_i64, ft, c_vp, c_i64, P_u32]),
    "sketch_dense_left": (c_int, [c_vp, c_i64, c_i64, c_int, c_int, c_i64, c_i64, c_i64, ft, c_i64, c_i64, c_vp, c_i64, ft, c_vp, c_i64,
                                  P_u32]),
    "sketch_dense_right": (c_int, [c_vp, c_i64, c_i64, c_int, c_int, c_i64, c_i64, c_i64, ft, c_vp, c_i64, c_i64, c_i64, ft, c_vp, c_i64,
                                   P_u32]),
}
SIGNATURES = {
    "rlb200_abi_version": (c_int, []),
    "rlb200_create": (c_int, [ctypes.POINTER(c_vp), c_int, c_vp]),
    "rlb200_destroy": (c_int, [c_vp]),
    "rlb200_last_error": (ctypes.c_char_p, [c_vp]),
    "rlb200_set_stream": (c_int, [c_vp, c_vp]),
    "rlb200_synchronize": (c_int, [c_vp]),
    "rlb200_set_shard": (c_int, [c_vp, c_i64, c_i64, ALLREDUCE_FN, c_vp]),
    "rlb200_launch_count": (c_i64, [c_vp, c_int]),
    "rlb200_timers_enable": (c_int, [c_vp, c_int]),
    "rlb200_timer_read": (c_int, [c_vp, c_int, ctypes.POINTER(ctypes.c_double), P_i64, c_int]),
    "rlb200_philox_stream_dev": (c_int, [c_vp, P_u32, c_i64, c_vp]),
    "rlb200_gemm_f64_i8_dev": (c_int, [c_vp, c_int, c_int, c_i64, c_i64, c_i64, ctypes.c_double, c_vp, c_i64, c_vp, c_i64, ctypes.c_double, c_vp,
                                       c_i64]),
    "rlb200_gemm_f32_i8_dev": (c_int, [c_vp, c_int, c_int, c_i64, c_i64, c_i64, ctypes.c_float, c_vp, c_i64, c_vp, c_i64, ctypes.c_float, c_vp,
                                       c_i64]),
    "rlb200_set_fp64_engine": (c_int, [c_vp, c_int]),
    "rlb200_set_i8_digits": (c_int, [c_vp, c_int]),
    "rlb200_set_i8_fused": (c_int, [c_vp, c_int]),
    "rlb200_set_shard_rank": (c_int, [c_vp, c_int, c_int]),
    "rlb200_set_phase_timing": (c_int, [c_vp, c_int]),
    "rlb200_set_bqrrp_tol": (c_int, [c_vp, ctypes.c_double]),
    "rlb200_set_cqrrpt_qrcp": (c_int, [c_vp, c_int]),
    "rlb200_set_cqrrpt_orthogonalization": (c_int, [c_vp, c_int]),
    "rlb200_set_cqrrpt_hqrrp_opts": (c_int, [c_vp, c_i64, c_i64, c_int, c_int]),
    "rlb200_get_phase_times": (c_int, [c_vp, c_vp, c_int]),
    "rlb200_comm_unique_id": (c_int, [c_vp]),
    "rlb200_comm_init": (c_int, [c_vp, c_int, c_int, c_vp]),
    "rlb200_comm_destroy": (c_int, [c_vp]),
    "rlb200_dev_alloc": (c_int, [c_vp, ctypes.c_size_t, ctypes.POINTER(c_vp)]),
    "rlb200_dev_free": (c_int, [c_vp, c_vp]),
    "rlb200_copy_h2d": (c_int, [c_vp, c_vp, c_vp, ctypes.c_size_t]),
    "rlb200_copy_d2h": (c_int, [c_vp, c_vp, c_vp, ctypes.c_size_t]),
}
for _suf, _ft in (("f64", ctypes.c_double), ("f32", ctypes.c_float)):
    for _name, _sig in _F(_ft).items():
        SIGNATURES[f"rlb200_{_name}_{_suf}_dev"] = _sig
    SIGNATURES[f"rlb200_cqrrpt_{_suf}_host"] = (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, _ft, _ft, c_i64, P_i64, P_u32])
    SIGNATURES[f"rlb200_bqrrp_{_suf}_host"] = (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, _ft, c_i64, c_int, c_int, c_vp, c_vp, P_i64, P_u32])
    SIGNATURES[f"rlb200_hqrrp_{_suf}_dev"] = (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_int, c_int, P_u32])
    SIGNATURES[f"rlb200_hqrrp_{_suf}_host"] = SIGNATURES[f"rlb200_hqrrp_{_suf}_dev"]
    SIGNATURES[f"rlb200_sketch_general_dense_left_{_suf}_dev"] = (c_int, [c_vp, c_int, c_int, c_int, c_i64, c_i64, c_i64, _ft, c_i64, c_i64, c_int, c_int,
                                                                         c_i64, c_i64, c_vp, c_i64, _ft, c_vp, c_i64, P_u32])
    SIGNATURES[f"rlb200_sketch_general_dense_right_{_suf}_dev"] = (c_int, [c_vp, c_int, c_int, c_int, c_i64, c_i64, c_i64, _ft, c_vp, c_i64, c_i64, c_i64,
                                                                          c_int, c_int, c_i64, c_i64, _ft, c_vp, c_i64, P_u32])
    SIGNATURES[f"rlb200_sketch_sparse_left_laso_{_suf}_dev"] = _F(_ft)["sketch_sparse_left"]
    SIGNATURES[f"rlb200_sketch_general_sparse_left_{_suf}_dev"] = (c_int, [c_vp, c_int, c_int, c_int, c_i64, c_i64, c_i64, _ft, c_i64, c_i64, c_i64,
                                                                          c_i64, c_i64, c_vp, c_i64, _ft, c_vp, c_i64, P_u32])
    SIGNATURES[f"rlb200_sketch_general_sparse_right_{_suf}_dev"] = (c_int, [c_vp, c_int, c_int, c_int, c_i64, c_i64, c_i64, _ft, c_vp, c_i64, c_i64, c_i64,
                                                                           c_i64, c_i64, c_i64, _ft, c_vp, c_i64, P_u32])
    SIGNATURES[f"rlb200_cqrrt_{_suf}_dev"] = (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, _ft, c_i64, c_int, c_int, P_u32])
    SIGNATURES[f"rlb200_cqrrt_{_suf}_host"] = (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, _ft, c_i64, c_int, c_int, P_u32])
    SIGNATURES[f"rlb200_syps_{_suf}_dev"] = (c_int, [c_vp, c_int, c_i64, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, P_u32])
    SIGNATURES[f"rlb200_syrf_{_suf}_dev"] = (c_int, [c_vp, c_int, c_i64, c_vp, c_i64, c_i64, c_vp, c_vp, P_u32, ctypes.POINTER(Revd2Opts)])
    SIGNATURES[f"rlb200_revd2_{_suf}_dev"] = (c_int, [c_vp, c_int, c_i64, c_vp, c_i64, P_i64, c_i64, _ft, c_vp, c_vp, P_u32,
                                                       ctypes.POINTER(Revd2Opts), ctypes.POINTER(_ft)])
    SIGNATURES[f"rlb200_revd2_{_suf}_host"] = SIGNATURES[f"rlb200_revd2_{_suf}_dev"]
    SIGNATURES[f"rlb200_bqrrp_{_suf}_dev_sk"] = (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, P_i64])
    SIGNATURES[f"rlb200_rsvd_{_suf}_host"] = (c_int, [c_vp, c_i64, c_i64, c_vp, P_i64, _ft, c_vp, c_vp, c_vp, P_u32,
                                                      ctypes.POINTER(StackOpts), P_int])

_lib = None


def load() -> ctypes.CDLL:
    """Load librlb200.so and declare every signature.  Raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C randlapack_b200/csrc). There is no fallback implementation.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib
