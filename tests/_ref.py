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


def _ft(dt):
    return ctypes.c_double if dt == np.float64 else ctypes.c_float


def _suf(dt):
    return "f64" if dt == np.float64 else "f32"


def ref_fill_sparse(lib, n_rows, n_cols, vec_nnz, seed6, dtype=np.float64, axis=1, sub=None, prefix="rlref"):
    """RandBLAS::fill_sparse_unpacked via the compiled reference -> (rc, nnz, vals, rows, cols, returned state)."""
    sr, sc, ro, co = sub if sub else (n_rows, n_cols, 0, 0)
    cap = vec_nnz * max(n_rows, n_cols)
    vals = np.zeros(cap, dtype=dtype)
    rows = np.full(cap, -1, dtype=np.int64)
    cols = np.full(cap, -1, dtype=np.int64)
    nnz = i64(0)
    st = (u32 * 6)(*seed6)
    f = getattr(lib, f"{prefix}_fill_sparse_{_suf(dtype)}")
    f.argtypes = [i64, i64, i64, ctypes.c_int, i64, i64, i64, i64, ctypes.POINTER(i64), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                  ctypes.POINTER(u32)]
    rc = f(n_rows, n_cols, vec_nnz, axis, sr, sc, ro, co, ctypes.byref(nnz), vals.ctypes.data, rows.ctypes.data, cols.ctypes.data, st)
    k = nnz.value
    return rc, k, vals[:k], rows[:k], cols[:k], list(st)


def ref_sketch_sparse_left(lib, S_rows, S_cols, vec_nnz, d, A, seed6, alpha=1.0, beta=0.0, B=None, ro=0, co=0):
    m, n = A.shape
    dt = A.dtype
    A = np.asfortranarray(A)
    B = np.zeros((d, n), dtype=dt, order="F") if B is None else np.asfortranarray(B.copy())
    st = (u32 * 6)(*seed6)
    f = getattr(lib, f"rlref_sketch_sparse_left_{_suf(dt)}")
    ft = _ft(dt)
    f.argtypes = [i64, i64, i64, i64, i64, i64, ft, i64, i64, ctypes.c_void_p, i64, ft, ctypes.c_void_p, i64, ctypes.POINTER(u32)]
    rc = f(S_rows, S_cols, vec_nnz, d, n, m, alpha, ro, co, A.ctypes.data, m, beta, B.ctypes.data, d, st)
    return rc, B, list(st)


def ref_sketch_dense(lib, left, S_rows, S_cols, d, A, seed6, family=0, axis=0, alpha=1.0, beta=0.0, B=None, ro=0, co=0):
    """left: B(d x n) = alpha S A + beta B with A m x n; right: B(m x d) = alpha A S + beta B."""
    m, n = A.shape
    dt = A.dtype
    A = np.asfortranarray(A)
    shape = (d, n) if left else (m, d)
    B = np.zeros(shape, dtype=dt, order="F") if B is None else np.asfortranarray(B.copy())
    st = (u32 * 6)(*seed6)
    ft = _ft(dt)
    if left:
        f = getattr(lib, f"rlref_sketch_dense_left_{_suf(dt)}")
        f.argtypes = [i64, i64, ctypes.c_int, ctypes.c_int, i64, i64, i64, ft, i64, i64, ctypes.c_void_p, i64, ft, ctypes.c_void_p, i64,
                      ctypes.POINTER(u32)]
        rc = f(S_rows, S_cols, family, axis, d, n, m, alpha, ro, co, A.ctypes.data, m, beta, B.ctypes.data, d, st)
    else:
        f = getattr(lib, f"rlref_sketch_dense_right_{_suf(dt)}")
        f.argtypes = [i64, i64, ctypes.c_int, ctypes.c_int, i64, i64, i64, ft, ctypes.c_void_p, i64, i64, i64, ft, ctypes.c_void_p, i64,
                      ctypes.POINTER(u32)]
        rc = f(S_rows, S_cols, family, axis, m, d, n, alpha, A.ctypes.data, m, ro, co, beta, B.ctypes.data, m, st)
    return rc, B, list(st)


def ref_cqrrpt(lib, A, d_factor, seed6, eps=None, nnz=2, qrcp=0, orthogonalization=False):
    """RandLAPACK::CQRRPT::call via the compiled reference -> (rc, rank, Q m x n [first rank cols valid], R n x n, J, state)."""
    m, n = A.shape
    dt = A.dtype
    Q = np.asfortranarray(A.copy())
    R = np.zeros((n, n), dtype=dt, order="F")
    J = np.zeros(n, dtype=np.int64)
    rank = i64(0)
    st = (u32 * 6)(*seed6)
    ft = _ft(dt)
    eps = float(np.finfo(dt).eps) ** 0.85 if eps is None else eps
    f = getattr(lib, f"rlref_cqrrpt_{'orth_' if orthogonalization else ''}{_suf(dt)}")
    f.argtypes = [i64, i64, ctypes.c_void_p, i64, ctypes.c_void_p, i64, ctypes.c_void_p, ft, ft, i64, ctypes.c_int, ctypes.POINTER(i64),
                  ctypes.POINTER(u32)]
    rc = f(m, n, Q.ctypes.data, m, R.ctypes.data, n, J.ctypes.data, d_factor, eps, nnz, qrcp, ctypes.byref(rank), st)
    return rc, rank.value, Q, R, J, list(st)


def ref_cqrrt(lib, A, d_factor, seed6, nnz=2, orthogonalization=False, compute_Q=True):
    """RandLAPACK::CQRRT::call via the compiled reference -> (rc, Q m x n, R n x n, state)."""
    m, n = A.shape
    dt = A.dtype
    Q = np.asfortranarray(A.copy())
    R = np.zeros((n, n), dtype=dt, order="F")
    st = (u32 * 6)(*seed6)
    ft = _ft(dt)
    f = getattr(lib, f"rlref_cqrrt_{_suf(dt)}")
    f.argtypes = [i64, i64, ctypes.c_void_p, i64, ctypes.c_void_p, i64, ft, ft, i64, ctypes.c_int, ctypes.c_int, ctypes.POINTER(u32)]
    rc = f(m, n, Q.ctypes.data, m, R.ctypes.data, n, d_factor, float(np.finfo(dt).eps), nnz, int(orthogonalization), int(compute_Q), st)
    return rc, Q, R, list(st)


def ref_syps(lib, uplo, A, k, p, q, seed6):
    """RandLAPACK::SYPS::call via the compiled reference -> (rc, skop m x k, state).  uplo: 0 upper / 1 lower."""
    m = A.shape[0]
    dt = A.dtype
    F = np.asfortranarray(A)
    sk = np.zeros((m, k), dtype=dt, order="F")
    st = (u32 * 6)(*seed6)
    f = getattr(lib, f"rlref_syps_{_suf(dt)}")
    f.argtypes = [ctypes.c_int, i64, ctypes.c_void_p, i64, i64, i64, i64, ctypes.c_void_p, ctypes.POINTER(u32)]
    rc = f(uplo, m, F.ctypes.data, m, k, p, q, sk.ctypes.data, st)
    return rc, sk, list(st)


def ref_syrf(lib, uplo, A, k, p, q, orth, seed6):
    """RandLAPACK::SYRF::call via the compiled reference -> (rc, Q m x k, state)."""
    m = A.shape[0]
    dt = A.dtype
    F = np.asfortranarray(A)
    Q = np.zeros((m, k), dtype=dt, order="F")
    st = (u32 * 6)(*seed6)
    f = getattr(lib, f"rlref_syrf_{_suf(dt)}")
    f.argtypes = [ctypes.c_int, i64, ctypes.c_void_p, i64, i64, i64, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(u32)]
    rc = f(uplo, m, F.ctypes.data, k, p, q, orth, Q.ctypes.data, st)
    return rc, Q, list(st)


def ref_revd2(lib, uplo, A, k, tol, p, q, orth, error_est_p, seed6, k_cap=None):
    """RandLAPACK::REVD2::call via the compiled reference -> (rc, k, V m x k, eigvals k, state)."""
    m = A.shape[0]
    dt = A.dtype
    k_cap = m if k_cap is None else k_cap
    F = np.asfortranarray(A)
    V = np.zeros((m, k_cap), dtype=dt, order="F")
    ev = np.zeros(k_cap, dtype=dt)
    kk = i64(k)
    st = (u32 * 6)(*seed6)
    f = getattr(lib, f"rlref_revd2_{_suf(dt)}")
    f.argtypes = [ctypes.c_int, i64, ctypes.c_void_p, ctypes.POINTER(i64), i64, _ft(dt), i64, i64, ctypes.c_int, ctypes.c_int,
                  ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(u32)]
    rc = f(uplo, m, F.ctypes.data, ctypes.byref(kk), k_cap, tol, p, q, orth, error_est_p, V.ctypes.data, ev.ctypes.data, st)
    return rc, kk.value, V[:, :kk.value].copy(order="F"), ev[:kk.value].copy(), list(st)


def ref_bqrrp(lib, A, d_factor, b_sz, seed6, qrcp_wide=0, qr_tall=0, tol=None):
    """RandLAPACK::BQRRP::call via the compiled reference -> (rc, rank, A_out [GEQP3 format], tau, J, state).  tol: the object's public field."""
    m, n = A.shape
    dt = A.dtype
    F = np.asfortranarray(A.copy())
    tau = np.zeros(n, dtype=dt)
    J = np.zeros(n, dtype=np.int64)
    rank = i64(0)
    st = (u32 * 6)(*seed6)
    if tol is not None:
        f = getattr(lib, f"rlref_bqrrp_tol_{_suf(dt)}")
        f.argtypes = [i64, i64, ctypes.c_void_p, i64, _ft(dt), i64, ctypes.c_int, ctypes.c_int, _ft(dt), ctypes.c_void_p, ctypes.c_void_p,
                      ctypes.POINTER(i64), ctypes.POINTER(u32)]
        rc = f(m, n, F.ctypes.data, m, d_factor, b_sz, qrcp_wide, qr_tall, tol, tau.ctypes.data, J.ctypes.data, ctypes.byref(rank), st)
        return rc, rank.value, F, tau, J, list(st)
    f = getattr(lib, f"rlref_bqrrp_{_suf(dt)}")
    f.argtypes = [i64, i64, ctypes.c_void_p, i64, _ft(dt), i64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                  ctypes.POINTER(i64), ctypes.POINTER(u32)]
    rc = f(m, n, F.ctypes.data, m, d_factor, b_sz, qrcp_wide, qr_tall, tau.ctypes.data, J.ctypes.data, ctypes.byref(rank), st)
    return rc, rank.value, F, tau, J, list(st)


def ref_hqrrp(lib, A, nb_alg, pp, panel_pivoting, qr_type, seed6):
    """RandLAPACK::hqrrp via the compiled reference -> (rc, A_out [GEQP3 format], tau, J, state)."""
    m, n = A.shape
    dt = A.dtype
    F = np.asfortranarray(A.copy())
    tau = np.zeros(n, dtype=dt)
    J = np.zeros(n, dtype=np.int64)
    st = (u32 * 6)(*seed6)
    f = getattr(lib, f"rlref_hqrrp_{_suf(dt)}")
    f.argtypes = [i64, i64, ctypes.c_void_p, i64, ctypes.c_void_p, ctypes.c_void_p, i64, i64, i64, i64, ctypes.POINTER(u32)]
    rc = f(m, n, F.ctypes.data, max(m, 1), J.ctypes.data, tau.ctypes.data, nb_alg, pp, panel_pivoting, qr_type, st)
    return rc, F, tau, J, list(st)


def ref_sketch_general_dense(lib, left, layout, opS, opA, D, dims, A_flat, lda, B_flat, ldb, seed6, alpha=1.0, beta=0.0, ro=0, co=0):
    """RandBLAS::sketch_general with every flag via the compiled reference.  D = (S_rows, S_cols, family, axis); dims = (d, n, m);
    layout: 1 ColMajor, 2 RowMajor; A_flat / B_flat: 1-D buffers holding the matrices in `layout` order with lda / ldb -> (rc, B_flat, state)."""
    dt = A_flat.dtype
    B = B_flat.copy()
    st = (u32 * 6)(*seed6)
    ft = _ft(dt)
    f = getattr(lib, f"rlref_sketch_general_dense_{_suf(dt)}")
    f.argtypes = [ctypes.c_int] * 4 + [i64, i64, ctypes.c_int, ctypes.c_int, i64, i64, i64, ft, i64, i64, ctypes.c_void_p, i64, ft, ctypes.c_void_p, i64,
                                       ctypes.POINTER(u32)]
    d, n, m = dims
    rc = f(int(left), layout, int(opS), int(opA), D[0], D[1], D[2], D[3], d, n, m, alpha, ro, co, A_flat.ctypes.data, lda, beta, B.ctypes.data, ldb, st)
    return rc, B, list(st)


def ref_sketch_general_sparse(lib, left, layout, opS, opA, D, dims, A_flat, lda, B_flat, ldb, seed6, alpha=1.0, beta=0.0, ro=0, co=0):
    """sketch_general with a short-axis SparseSkOp and every flag via the compiled reference.  D = (S_rows, S_cols, vec_nnz); dims = (d, n, m)."""
    dt = A_flat.dtype
    B = B_flat.copy()
    st = (u32 * 6)(*seed6)
    ft = _ft(dt)
    f = getattr(lib, f"rlref_sketch_general_sparse_axis_{_suf(dt)}")
    f.argtypes = [ctypes.c_int] * 4 + [i64, i64, i64, ctypes.c_int, i64, i64, i64, ft, i64, i64, ctypes.c_void_p, i64, ft, ctypes.c_void_p, i64,
                                       ctypes.POINTER(u32)]
    d, n, m = dims
    axis = D[3] if len(D) > 3 else 1          # RL_AXIS_SHORT
    rc = f(int(left), layout, int(opS), int(opA), D[0], D[1], D[2], axis, d, n, m, alpha, ro, co, A_flat.ctypes.data, lda, beta, B.ctypes.data, ldb, st)
    return rc, B, list(st)
