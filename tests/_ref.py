"""Test helpers: load the two CPU checkers (oracle restatement, compiled reference) via ctypes."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
u32, i64 = ctypes.c_uint32, ctypes.c_int64


class Opts(ctypes.Structure):
    _fields_ = [("p", i64), ("q", i64), ("b", i64), ("stab", ctypes.c_int32), ("orth_rf", ctypes.c_int32),
                ("orth_qb", ctypes.c_int32), ("cond", ctypes.c_int32), ("orthc", ctypes.c_int32), ("res", ctypes.c_int32)]


def ref_lib():
    p = os.path.join(ROOT, "oracle", "_ref", "librl_ref.so")
    return ctypes.CDLL(p) if os.path.exists(p) else None


def oracle_lib():
    return ctypes.CDLL(os.path.join(ROOT, "oracle", "librl_oracle.so"))


def fill_dense(lib, prefix, n_rows, n_cols, seed6, dtype=np.float64, family=0, axis=0, layout=0, sub=None):
    sr, sc, ro, co = sub if sub else (n_rows, n_cols, 0, 0)
    st = (u32 * 6)(*seed6)
    buf = np.full(sr * sc, -7, dtype=dtype)
    f = getattr(lib, f"{prefix}_fill_dense_{'f64' if dtype == np.float64 else 'f32'}")
    f.argtypes = [i64, i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, i64, i64, i64, i64, ctypes.c_void_p, ctypes.POINTER(u32)]
    rc = f(n_rows, n_cols, family, axis, layout, sr, sc, ro, co, buf.ctypes.data, st)
    return rc, buf, list(st)


def ref_mat_gen(lib, mtype, m, n, rank, cond, expo, seed6, dtype=np.float64):
    A = np.zeros((m, n), dtype=dtype, order="F")
    st = (u32 * 6)(*seed6)
    ft = ctypes.c_double if dtype == np.float64 else ctypes.c_float
    f = lib.rlref_mat_gen_f64 if dtype == np.float64 else lib.rlref_mat_gen_f32
    f.argtypes = [ctypes.c_int, i64, i64, i64, ft, ft, ft, ctypes.c_void_p, ctypes.POINTER(u32)]
    rc = f(mtype, m, n, rank, cond, expo, 1.0, A.ctypes.data, st)
    assert rc == 0
    return A, list(st)


def ref_rsvd(lib, A, k, tol, seed6, o):
    m, n = A.shape
    dt = A.dtype
    U = np.zeros((m, k), dtype=dt, order="F")
    S = np.zeros(k, dtype=dt)
    V = np.zeros((n, k), dtype=dt, order="F")
    kk = i64(k)
    st = (u32 * 6)(*seed6)
    A = A.copy(order="F")
    ft = ctypes.c_double if dt == np.float64 else ctypes.c_float
    f = lib.rlref_rsvd_f64 if dt == np.float64 else lib.rlref_rsvd_f32
    f.argtypes = [i64, i64, ctypes.c_void_p, ctypes.POINTER(i64), ft, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                  ctypes.POINTER(u32), ctypes.POINTER(Opts)]
    oo = Opts(o.passes_over_data, o.passes_per_stab, o.block_sz, o.stab, o.orth_rf, o.orth_qb, int(o.cond_check), int(o.orth_check), 0)
    rc = f(m, n, A.ctypes.data, ctypes.byref(kk), tol, U.ctypes.data, S.ctypes.data, V.ctypes.data, st, ctypes.byref(oo))
    return rc, kk.value, U[:, :kk.value], S[:kk.value], V[:, :kk.value], list(st)


def ulp_diff_f32(a, b):
    """Distance in float32 ulps between two arrays holding float32-representable values."""
    a32, b32 = np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32)
    ia, ib = a32.view(np.int32).astype(np.int64), b32.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


def subspace_sin(Q1, Q2):
    """sin of the largest principal angle between range(Q1) and range(Q2) (orthonormal columns)."""
    M = Q2 - Q1 @ (Q1.T @ Q2)
    return np.linalg.norm(M, 2)
