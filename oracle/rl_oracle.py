"""CPU restatement of the RandLAPACK sketch-and-factor drivers (RS / RF / QB / RSVD and the
CholQRQ / PLUL / HQRQ stabilisers) on numpy + the LAPACK that scipy bundles.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this module; the product (randlapack_b200/) never does.

Every function cites the reference lines it follows (paths relative to the reference root).
The random-number layer (Philox4x32-10, Box-Muller, DenseDist counter layout) lives in
oracle/rl_oracle.c so that it uses the same libm calls as the reference's host path; this module
loads it through ctypes.  Parity pins: see the header of oracle/rl_oracle.c; the drivers here are
validated bit-for-bit / to round-off against the real reference compiled in oracle/_ref
(tests/test_oracle_drivers.py) and against the golden fixtures in tests/golden/.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np
import scipy.linalg
from scipy.linalg import get_blas_funcs, get_lapack_funcs

_HERE = os.path.dirname(os.path.abspath(__file__))

STAB_PLUL, STAB_CHOLQRQ, STAB_HQRQ = 0, 1, 2
FAMILY_GAUSSIAN, FAMILY_UNIFORM = 0, 1
AXIS_LONG, AXIS_SHORT = 0, 1
LAYOUT_NATURAL, LAYOUT_COLMAJOR, LAYOUT_ROWMAJOR = 0, 1, 2

_u32 = ctypes.c_uint32
_i64 = ctypes.c_int64
_lib = None


def lib() -> ctypes.CDLL:
    """Load oracle/librl_oracle.so (built by `make -C oracle oracle`)."""
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(os.path.join(_HERE, "librl_oracle.so"))
    return _lib


# --------------------------------------------------------------------------------------------
# RNG layer (thin ctypes veneer over rl_oracle.c)
# --------------------------------------------------------------------------------------------
class RNGState:
    """RandBLAS::RNGState<Philox4x32> (RandBLAS/RandBLAS/base.hh:64-164): 128-bit counter + 64-bit key."""

    def __init__(self, key: int = 0, counter=(0, 0, 0, 0)):
        if isinstance(key, (tuple, list)):
            k = tuple(int(x) & 0xFFFFFFFF for x in key)
        else:  # RNGState(uint64 k): key.incr(k) on a zero key (base.hh:119)
            k = (int(key) & 0xFFFFFFFF, (int(key) >> 32) & 0xFFFFFFFF)
        self.counter = tuple(int(c) & 0xFFFFFFFF for c in counter)
        self.key = k

    def words(self):
        return (_u32 * 6)(*self.counter, *self.key)

    @classmethod
    def from_words(cls, w):
        return cls(key=(w[4], w[5]), counter=(w[0], w[1], w[2], w[3]))

    def copy(self):
        return RNGState(self.key, self.counter)

    def __eq__(self, other):
        return self.counter == other.counter and self.key == other.key

    def __repr__(self):
        return f"RNGState(counter={self.counter}, key={self.key})"


def philox4x32_10(ctr, key):
    out = (_u32 * 4)()
    lib().rlo_philox4x32_10((_u32 * 4)(*ctr), (_u32 * 2)(*key), out)
    return tuple(out)


def ctr_incr(ctr, step: int):
    c = (_u32 * 4)(*ctr)
    lib().rlo_ctr_incr(c, ctypes.c_uint64(step))
    return tuple(c)


def fill_dense(n_rows, n_cols, state: RNGState, dtype=np.float64, family=FAMILY_GAUSSIAN, major_axis=AXIS_LONG,
               layout=LAYOUT_NATURAL, sub=None):
    """RandBLAS::fill_dense / fill_dense_unpacked (dense_skops.hh:560-603, 620-623).

    Returns (matrix as a 2-D numpy array in the requested layout, next RNGState)."""
    sub_rows, sub_cols, ro, co = sub if sub is not None else (n_rows, n_cols, 0, 0)
    dt = np.dtype(dtype)
    buf = np.empty(sub_rows * sub_cols, dtype=dt)
    w = state.words()
    fn = lib().rlo_fill_dense_f64 if dt == np.float64 else lib().rlo_fill_dense_f32
    fn.argtypes = [_i64, _i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, _i64, _i64, _i64, _i64, ctypes.c_void_p,
                   ctypes.POINTER(_u32)]
    rc = fn(n_rows, n_cols, family, major_axis, layout, sub_rows, sub_cols, ro, co, buf.ctypes.data, w)
    if rc:
        raise ValueError("fill_dense: invalid arguments (randblas_require failed)")
    is_wide, fa_long = n_rows < n_cols, major_axis == AXIS_LONG
    nat_col = (not is_wide and fa_long) or (is_wide and not fa_long)  # dense_skops.hh:184-196
    col = nat_col if layout == LAYOUT_NATURAL else layout == LAYOUT_COLMAJOR
    mat = buf.reshape((sub_rows, sub_cols), order="F" if col else "C")
    return mat, RNGState.from_words(w)


# --------------------------------------------------------------------------------------------
# BLAS/LAPACK veneer: same routines, same argument order as the reference's calls
# --------------------------------------------------------------------------------------------
def _F(a):
    return np.asfortranarray(a)


def _gemm(a, b, ta=False, tb=False, alpha=1.0, beta=0.0, c=None):
    (gemm,) = get_blas_funcs(("gemm",), (a, b))
    if c is None:
        return gemm(alpha, a, b, trans_a=int(ta), trans_b=int(tb))
    return gemm(alpha, a, b, beta=beta, c=c, trans_a=int(ta), trans_b=int(tb), overwrite_c=1)


# --------------------------------------------------------------------------------------------
# Stabilisers (RandLAPACK/comps/rl_orth.hh)
# --------------------------------------------------------------------------------------------
class CholQRQ:
    """rl_orth.hh:25-98 — syrk(Upper,Trans) -> potrf(Upper) -> [cond check] -> trsm(Right,Upper,NoTrans)."""

    def __init__(self, cond_check=False, verbose=False):
        self.cond_check, self.verbose, self.chol_fail = cond_check, verbose, False

    def call(self, A):
        m, k = A.shape
        syrk, trsm = get_blas_funcs(("syrk", "trsm"), (A,))
        (potrf,) = get_lapack_funcs(("potrf",), (A,))
        G = syrk(1.0, A, trans=1, lower=0)                       # :78
        R, info = potrf(G, lower=0, clean=0)                      # :81
        if info != 0:
            self.chol_fail = True
            return 1, A
        if self.cond_check:                                       # :88-93
            if cond_num(R) > 1.0 / np.sqrt(np.finfo(A.dtype).eps):
                return 1, A
        Q = trsm(1.0, R, A, side=1, lower=0, trans_a=0, diag=0)   # :95
        return 0, _F(Q)


class HQRQ:
    """rl_orth.hh:100-164 — geqrf + ungqr."""

    def __init__(self, cond_check=False, verbose=False):
        self.cond_check, self.verbose = cond_check, verbose

    def call(self, A):
        geqrf, orgqr = get_lapack_funcs(("geqrf", "orgqr"), (A,))
        qr, tau, _, info = geqrf(A)
        if info:
            return 1, A
        q, _, info = orgqr(qr, tau)
        return 0, _F(q)


class PLUL:
    """rl_orth.hh:166-230 — getrf (singular U tolerated) -> unit-lower L (get_L) -> laswp(1..n, incx=+1)."""

    def __init__(self, cond_check=False, verbose=False):
        self.cond_check, self.verbose = cond_check, verbose

    def call(self, A):
        m, n = A.shape
        (getrf, laswp) = get_lapack_funcs(("getrf", "laswp"), (A,))
        lu, piv, info = getrf(A)                                  # :222
        L = np.tril(lu, -1)                                       # util::get_L(m,n,A,1): rl_util.hh:101-114
        idx = np.arange(min(m, n))
        L[idx, idx] = 1.0
        L = _F(L)
        # lapack::laswp(n, A, m, 1, n, ipiv, 1) (:225): forward application of the interchanges
        L = laswp(L, piv, k1=0, k2=min(m, n) - 1, off=0, inc=1)
        return 0, _F(L)


def make_stab(kind, cond_check=False):
    return {STAB_PLUL: PLUL, STAB_CHOLQRQ: CholQRQ, STAB_HQRQ: HQRQ}[kind](cond_check, False)


def cond_num(A):
    """util::cond_num_check (rl_util.hh:402-424): gesdd(NoVec), s[0]/s[n-1] (inf if s[n-1]==0)."""
    (gesdd,) = get_lapack_funcs(("gesdd",), (A,))
    _, s, _, _ = gesdd(A, compute_uv=0)
    return np.inf if s[-1] == 0 else s[0] / s[-1]


def orthogonality_check(Q):
    """util::orthogonality_check (rl_util.hh:467-496): ||triu(Q'Q) - I||_F / sqrt(k) > tol."""
    k = Q.shape[1]
    (syrk,) = get_blas_funcs(("syrk",), (Q,))
    G = syrk(1.0, Q, trans=1, lower=0)
    G[np.arange(k), np.arange(k)] -= 1.0
    tol = 1e-10 if Q.dtype == np.float64 else 1e-2
    return np.linalg.norm(G) / np.sqrt(k) > tol


# --------------------------------------------------------------------------------------------
# RS / RF / QB / RSVD
# --------------------------------------------------------------------------------------------
@dataclass
class StackOpts:
    """Mirror of rl_stack_opts (oracle/oracle_capi.h) = the canonical stack of test/drivers/test_rsvd.cc:68-93."""
    passes_over_data: int = 0
    passes_per_stab: int = 1
    block_sz: int = 0
    stab: int = STAB_PLUL
    orth_rf: int = STAB_CHOLQRQ
    orth_qb: int = STAB_CHOLQRQ
    cond_check: bool = False
    orth_check: bool = False


class RS:
    """RandLAPACK::RS (rl_rs.hh:31-178)."""

    def __init__(self, stab, p, q, verbose=False, cond_check=False):
        self.stab, self.passes_over_data, self.passes_per_stab = stab, p, q
        self.cond_check, self.cond_nums = cond_check, []

    def call(self, A, k, state: RNGState, omega_override=None):
        """Returns (rc, Omega n-by-k, next_state). `omega_override` lets a parity test inject the very
        operator the device generated (cf. test/drivers/test_bqrrp_gpu.cu:91-103, where the reference
        feeds one host-built sketch to both its CPU and GPU paths)."""
        m, n = A.shape
        p, q, p_done = self.passes_over_data, self.passes_per_stab, 0
        dt = A.dtype
        if p % 2 == 0:                                             # :132-135
            Om, state = fill_dense(n, k, state, dt)
            # the natural-layout buffer is handed to BLAS as a column-major n x k matrix (:153), whatever D.natural_layout is
            Om = _F(np.ravel(Om, order="K").reshape((n, k), order="F"))
            if omega_override is not None:
                Om = _F(omega_override.astype(dt))
        else:                                                      # :136-149
            Om1, state = fill_dense(m, k, state, dt)
            Om1 = _F(np.ravel(Om1, order="K").reshape((m, k), order="F"))
            if omega_override is not None:
                Om1 = _F(omega_override.astype(dt))
            Om = _gemm(A, Om1, ta=True)
            p_done += 1
            if p_done % q == 0:
                rc, Om = self.stab.call(Om)
                if rc:
                    return 1, Om, state
        while p - p_done > 0:                                      # :151-174
            Om1 = _gemm(A, Om)
            p_done += 1
            if self.cond_check:
                self.cond_nums.append(cond_num(Om1))
            if p_done % q == 0:
                rc, Om1 = self.stab.call(Om1)
                if rc:
                    return 1, Om, state
            Om = _gemm(A, Om1, ta=True)
            p_done += 1
            if self.cond_check:
                self.cond_nums.append(cond_num(Om))
            if p_done % q == 0:
                rc, Om = self.stab.call(Om)
                if rc:
                    return 1, Om, state
        return 0, Om, state


class RF:
    """RandLAPACK::RF (rl_rf.hh:31-137)."""

    def __init__(self, rs, orth, verbose=False, cond_check=False):
        self.rs, self.orth, self.cond_check, self.cond_nums = rs, orth, cond_check, []

    def call(self, A, k, state, omega_override=None):
        rc, Om, state = self.rs.call(A, k, state, omega_override)   # :118
        if rc:
            return 1, None, state
        Q = _gemm(A, Om)                                             # :123
        if self.cond_check:
            self.cond_nums.append(cond_num(Q))
        rc, Q = self.orth.call(Q)                                    # :129
        if rc:
            return 2, Q, state
        return 0, Q, state


class QB:
    """RandLAPACK::QB (rl_qb.hh:36-268). Returns (rc, k, Q m-by-k, BT n-by-k, state)."""

    def __init__(self, rf, orth, verbose=False, orth_check=False):
        self.rf, self.orth, self.orth_check = rf, orth, orth_check

    def call(self, A, k, b_sz, tol, state, omega_override=None):
        m, n = A.shape
        dt = A.dtype
        eps = np.finfo(dt).eps
        tol = max(tol, 100 * eps)                                    # :149
        curr, norm_B, approx_err = 0, dt.type(0), dt.type(0)
        (lange,) = get_lapack_funcs(("lange",), (A,))
        norm_A = dt.type(lange("F", A))                              # :168
        A_cpy = _F(A.copy())                                         # :171
        Q = np.zeros((m, 0), dtype=dt, order="F")
        BT = np.zeros((n, 0), dtype=dt, order="F")
        while curr < k:
            b_sz = min(b_sz, k - curr)
            nxt = curr + b_sz
            rc, Q_i, state = self.rf.call(A_cpy, b_sz, state, omega_override)   # :191
            if rc:
                return 6, curr, Q, BT, state
            if self.orth_check and orthogonality_check(Q_i):         # :199-207
                return 4, curr, Q, BT, state
            if curr != 0:                                            # :210-215
                QtQi = _gemm(Q, Q_i, ta=True)
                Q_i = _gemm(Q, QtQi, alpha=-1.0, beta=1.0, c=_F(Q_i))
                _, Q_i = self.orth.call(Q_i)
            BT_i = _gemm(A_cpy, Q_i, ta=True)                        # :218
            norm_B_i = dt.type(lange("F", BT_i))
            norm_B = dt.type(np.hypot(norm_B, norm_B_i))             # :222
            prev_err = approx_err
            approx_err = dt.type(np.sqrt(np.abs(norm_A - norm_B)) * (np.sqrt(norm_A + norm_B) / norm_A))   # :225
            Qn = _F(np.hstack([Q, Q_i]))
            BTn = _F(np.hstack([BT, BT_i]))
            if curr > 0 and approx_err > prev_err:                   # :228-234 (Q,BT buffers already hold block i)
                return 2, curr, Qn, BTn, state
            if self.orth_check and orthogonality_check(Qn):          # :236-244
                return 5, curr, Qn, BTn, state
            Q, BT = Qn, BTn
            curr += b_sz
            if approx_err < tol:                                     # :250-256
                return 0, curr, Q, BT, state
            A_cpy = _gemm(Q_i, BT_i, tb=True, alpha=-1.0, beta=1.0, c=A_cpy)   # :260
        return 3, curr, Q, BT, state


class RSVD:
    """RandLAPACK::RSVD (rl_rsvd.hh:34-154). Returns (rc, k, U m-by-k, S k, V n-by-k, state)."""

    def __init__(self, qb, block_sz):
        self.qb, self.block_sz = qb, block_sz

    def call(self, A, k, tol, state, omega_override=None):
        m, n = A.shape
        if k <= 0 or tol < 0:                                        # :128-132
            raise ValueError("RandLAPACK::Error: invalid argument")
        _, k, Q, BT, state = self.qb.call(A, k, self.block_sz, tol, state, omega_override)   # :137 (rc ignored)
        (gesdd,) = get_lapack_funcs(("gesdd",), (BT,))
        # gesdd(SomeVec, n, k, BT, n, S, V, n, UT_buf, k): BT = V * diag(S) * UT_buf   (:146)
        V, S, UT, info = gesdd(_F(BT[:, :k]), compute_uv=1, full_matrices=0)
        U = _gemm(_F(Q[:, :k]), UT, tb=True)                         # :148
        return 0, k, _F(U), S, _F(V), state


def make_stack(o: StackOpts):
    stab = make_stab(o.stab, o.cond_check)
    rs = RS(stab, o.passes_over_data, o.passes_per_stab, False, o.cond_check)
    rf = RF(rs, make_stab(o.orth_rf, o.cond_check), False, o.cond_check)
    qb = QB(rf, make_stab(o.orth_qb, o.cond_check), False, o.orth_check)
    return rs, rf, qb, RSVD(qb, o.block_sz)


# --------------------------------------------------------------------------------------------
# Test-matrix generators the reference's tests use as inputs (RandLAPACK/testing/rl_gen.hh)
# --------------------------------------------------------------------------------------------
def gen_poly_singvals(k, frac_spectrum_one, cond, p, dtype=np.float64):
    """rl_gen.hh:105-126."""
    T = np.dtype(dtype).type
    s = np.empty(k, dtype=dtype)
    offset = int(np.floor(k * frac_spectrum_one))
    first, last = T(1.0), T(1.0) / T(cond)
    neg_invp = -T(1.0) / T(p)
    a = np.power((np.power(last, neg_invp) - np.power(first, neg_invp)) / T(k - offset), T(p))
    b = np.power(a * first, neg_invp) - offset
    s[:offset] = 1.0
    for i in range(offset, k):
        s[i] = 1 / (a * np.power(T(i) + b, T(p)))
    return s


def gen_singvec(m, n, S_diag, state, dtype=np.float64):
    """rl_gen.hh:62-92: A = Q_U diag(S) Q_V', Q_U/Q_V from Householder QR of Gaussian m-by-k / n-by-k."""
    k = len(S_diag)
    A = np.zeros((m, n), dtype=dtype, order="F")
    U, state = fill_dense(m, k, state, dtype)
    V, state = fill_dense(n, k, state, dtype)
    U, V = _F(U), _F(V)
    A[np.arange(k), np.arange(k)] = S_diag
    geqrf, ormqr = get_lapack_funcs(("geqrf", "ormqr"), (A,))

    def _ormqr(side, trans, qr, tau, c):
        _, lw, _ = ormqr(side, trans, qr, tau, c, lwork=-1)
        out, _, info = ormqr(side, trans, qr, tau, c, lwork=int(lw[0]))
        return _F(out)
    qr, tau, _, _ = geqrf(U)
    A = _ormqr("L", "N", qr, tau, A)
    qr, tau, _, _ = geqrf(V)
    A = _ormqr("R", "T", qr, tau, A)
    return A, state


def gen_poly_mat(m, n, k, cond, exponent, state, frac_spectrum_one=0.1, dtype=np.float64):
    """rl_gen.hh gen_poly_mat (mat_gen case `polynomial`, :720-723) with diag=false."""
    s = gen_poly_singvals(k, frac_spectrum_one, cond, exponent, dtype)
    return gen_singvec(m, n, s, state, dtype)


def gen_adversarial_mat(m, n, sigma, state, dtype=np.float64):
    """rl_gen.hh:311-359 (mat_gen case `adverserial`): A = orth(G_U with 10 rows scaled by sigma) * triu(orth(G_V)) with the
    diagonal of the triangular factor scaled by 1e-2 from column 11 on."""
    U, state = fill_dense(m, n, state, dtype)
    V, state = fill_dense(n, n, state, dtype)
    U, V = _F(U.copy()), _F(V.copy())
    U[:10, :] *= np.dtype(dtype).type(sigma)
    geqrf, orgqr = get_lapack_funcs(("geqrf", "orgqr"), (U,))

    def _q(M):
        qr, tau, _, _ = geqrf(M)
        _, lw, _ = orgqr(qr, tau, lwork=-1)
        q, _, info = orgqr(qr, tau, lwork=int(lw[0]))
        return _F(q)
    U, V = _q(U), np.triu(_q(V))
    for i in range(11, n):
        V[i, i] *= np.dtype(dtype).type(10e-3)
    return _gemm(U, _F(V)), state


# --------------------------------------------------------------------------------------------
# RandBLAS sketching operators applied (sparse_skops.hh, skge.hh)
# --------------------------------------------------------------------------------------------
def fill_sparse(n_rows, n_cols, vec_nnz, state: RNGState, dtype=np.float64, sub=None):
    """RandBLAS::fill_sparse_unpacked for SparseDist(n_rows, n_cols, vec_nnz, Axis::Short) (sparse_skops.hh:568-704).
    -> (nnz, vals, rows, cols, returned RNGState)."""
    sr, sc, ro, co = sub if sub is not None else (n_rows, n_cols, 0, 0)
    cap = max(1, vec_nnz * max(sr, sc))
    rows, cols = np.empty(cap, dtype=np.int64), np.empty(cap, dtype=np.int64)
    vals = np.empty(cap, dtype=np.float64)
    nnz = _i64(0)
    w = state.words()
    fn = lib().rlo_saso_coo
    fn.argtypes = [_i64] * 7 + [ctypes.POINTER(_i64), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(_u32)]
    rc = fn(n_rows, n_cols, vec_nnz, sr, sc, ro, co, ctypes.byref(nnz), rows.ctypes.data, cols.ctypes.data, vals.ctypes.data, w)
    if rc:
        raise ValueError("fill_sparse: invalid arguments (randblas_require failed)")
    k = nnz.value
    return k, vals[:k].astype(dtype), rows[:k].copy(), cols[:k].copy(), RNGState.from_words(w)


def fill_sparse_laso(n_rows, n_cols, vec_nnz, state: RNGState, dtype=np.float64, sub=None):
    """RandBLAS::fill_sparse_unpacked for SparseDist(n_rows, n_cols, vec_nnz, Axis::Long) (sparse_skops.hh:585-610, 669-704): every long-axis
    vector draws vec_nnz iid uniform indices (one Philox counter each: (rv0 + 2^32 rv1) mod dim_major, sign from rv2, util.hh:521-541),
    duplicates are merged into sqrt(count) * first sign (:483-511), survivors sorted by index, then filtered to the sub-matrix window.
    -> (nnz, vals, rows, cols, returned RNGState)."""
    sr, sc, ro, co = sub if sub is not None else (n_rows, n_cols, 0, 0)
    dt = np.dtype(dtype)
    short_rows = n_rows <= n_cols
    dim_major = max(n_rows, n_cols)
    vec_off, vec_sub = (ro, sr) if short_rows else (co, sc)
    lo, ls = (co, sc) if short_rows else (ro, sr)
    ctr = ctr_incr(state.counter, vec_off * vec_nnz)
    maj, mino, vals = [], [], []
    for v in range(vec_sub):
        first, count = {}, {}
        for _ in range(vec_nnz):
            rv = philox4x32_10(ctr, state.key)
            ctr = ctr_incr(ctr, 1)
            ell = (rv[0] + (rv[1] << 32)) % dim_major
            if ell in count:
                count[ell] += 1
            else:
                count[ell] = 1
                first[ell] = 1.0 if rv[2] % 2 == 0 else -1.0
        for ell in sorted(count):
            if lo <= ell < lo + ls:
                maj.append(ell - lo)
                mino.append(v)
                vals.append(np.sqrt(dt.type(count[ell])) * dt.type(first[ell]))
    maj, mino = np.array(maj, dtype=np.int64), np.array(mino, dtype=np.int64)
    rows, cols = (mino, maj) if short_rows else (maj, mino)
    return len(vals), np.array(vals, dtype=dt), rows, cols, RNGState(state.key, ctr)


def laso_next_state(n_rows, n_cols, vec_nnz, state: RNGState):
    """compute_next_state for Axis::Long (sparse_skops.hh:302-312): min(n_rows, n_cols) vectors of vec_nnz draws."""
    return RNGState(state.key, ctr_incr(state.counter, min(n_rows, n_cols) * vec_nnz))


def saso_next_state(n_rows, n_cols, vec_nnz, state: RNGState):
    """SparseSkOp::next_state (sparse_skops.hh:302-312)."""
    w = state.words()
    lib().rlo_saso_next_state(_i64(n_rows), _i64(n_cols), _i64(vec_nnz), w)
    return RNGState.from_words(w)


def sketch_sparse_left(S_rows, S_cols, vec_nnz, d, A, state: RNGState, alpha=1.0, beta=0.0, B=None, ro=0, co=0):
    """sketch_general(ColMajor, NoTrans, NoTrans, d, n, m, alpha, S, ro, co, A, lda, beta, B, ldb), S a wide SASO
    (skge.hh:538-571; the +-1 entries are NOT scaled by isometry_scale).  -> (B, S.next_state)."""
    from scipy.sparse import coo_matrix
    m, n = A.shape
    k, vals, rows, cols, _ = fill_sparse(S_rows, S_cols, vec_nnz, state, A.dtype, sub=(d, m, ro, co))
    S = coo_matrix((vals, (rows, cols)), shape=(d, m)).tocsr()
    out = alpha * (S @ A)
    if B is not None and beta != 0:
        out = out + beta * B
    return _F(out.astype(A.dtype)), saso_next_state(S_rows, S_cols, vec_nnz, state)


def dense_next_state(n_rows, n_cols, major_axis, state: RNGState):
    w = state.words()
    lib().rlo_dense_next_state(_i64(n_rows), _i64(n_cols), ctypes.c_int(major_axis), w)
    return RNGState.from_words(w)


def sketch_dense_left(S_rows, S_cols, d, A, state: RNGState, family=FAMILY_GAUSSIAN, major_axis=AXIS_LONG, alpha=1.0, beta=0.0,
                      B=None, ro=0, co=0):
    """lskge3 (skge.hh:155-203): B = alpha * S[ro:ro+d, co:co+m] A + beta B.  -> (B, S.next_state)."""
    m, n = A.shape
    S, _ = fill_dense(S_rows, S_cols, state, A.dtype, family, major_axis, LAYOUT_COLMAJOR, sub=(d, m, ro, co))
    out = alpha * (S @ A)
    if B is not None and beta != 0:
        out = out + beta * B
    return _F(out), dense_next_state(S_rows, S_cols, major_axis, state)


def sketch_dense_right(A, S_rows, S_cols, d, state: RNGState, family=FAMILY_GAUSSIAN, major_axis=AXIS_LONG, alpha=1.0, beta=0.0,
                       B=None, ro=0, co=0):
    """rskge3 (skge.hh:308-356): B = alpha * A S[ro:ro+n, co:co+d] + beta B."""
    m, n = A.shape
    S, _ = fill_dense(S_rows, S_cols, state, A.dtype, family, major_axis, LAYOUT_COLMAJOR, sub=(n, d, ro, co))
    out = alpha * (A @ S)
    if B is not None and beta != 0:
        out = out + beta * B
    return _F(out), dense_next_state(S_rows, S_cols, major_axis, state)


# --------------------------------------------------------------------------------------------
# CQRRPT (RandLAPACK/drivers/rl_cqrrpt.hh:146-391), default subroutines (SASO sketch, geqp3)
# --------------------------------------------------------------------------------------------
def _mat_view(flat, rows, cols, ld, layout):
    """rows x cols view into a flat buffer stored in `layout` (LAYOUT_COLMAJOR | LAYOUT_ROWMAJOR) with leading dimension ld."""
    if layout == LAYOUT_COLMAJOR:
        return flat[:ld * cols].reshape((ld, cols), order="F")[:rows, :]
    return flat[:rows * ld].reshape((rows, ld))[:, :cols]


def sketch_general_dense(left, layout, opS, opA, d, n, m, alpha, S_rows, S_cols, ro, co, A_flat, lda, beta, B_flat, ldb, state: RNGState,
                         family=FAMILY_GAUSSIAN, major_axis=AXIS_LONG):
    """RandBLAS::sketch_general with a DenseSkOp and every layout / transposition flag (skge.hh:859-905 left -> lskge3 :100-203;
    :1031-1076 right -> rskge3 :253-356), restated on the materialised operator.
    left : B(d x n) = alpha * op(S[ro:, co:])(d x m) * op(A)(m x n) + beta * B;   right: B(m x d) = alpha * op(A)(m x n) * op(S[ro:, co:])(n x d) + beta * B.
    A_flat / B_flat: flat buffers in `layout` order.  -> (B_flat (new), S.next_state)."""
    dt = A_flat.dtype
    S, nxt = fill_dense(S_rows, S_cols, state, dt, family, major_axis)
    rs, cs = ((m, d) if opS else (d, m)) if left else ((d, n) if opS else (n, d))           # dims_before_op
    if S_rows < rs + ro or S_cols < cs + co:
        raise ValueError("sketch_general: submatrix of S out of range (randblas_require)")
    Sub = np.asarray(S)[ro:ro + rs, co:co + cs]
    opSm = Sub.T if opS else Sub
    ra, ca = (n, m) if opA else (m, n)
    A2 = _mat_view(A_flat, ra, ca, lda, layout)
    opAm = A2.T if opA else A2
    R = (opSm @ opAm) if left else (opAm @ opSm)
    out = np.array(B_flat, copy=True)
    Bv = _mat_view(out, *((d, n) if left else (m, d)), ldb, layout)
    Bv[:, :] = dt.type(alpha) * R + (dt.type(beta) * Bv if beta != 0 else 0)
    return out, nxt


def sketch_general_sparse(left, layout, opS, opA, d, n, m, alpha, S_rows, S_cols, vec_nnz, ro, co, A_flat, lda, beta, B_flat, ldb, state: RNGState,
                          major_axis=AXIS_SHORT):
    """RandBLAS::sketch_general with a short-axis SparseSkOp and every layout / transposition flag (skge.hh:907-960 left, :1078-1131 right),
    restated on the materialised operator.  A tall short-axis operator is generated as the transpose of the wide one with swapped dimensions
    and the same seed (sparse_skops.hh:585-610: the index stream depends on (dim_major, dim_minor) only).  -> (B_flat (new), S.next_state)."""
    dt = A_flat.dtype
    if major_axis == AXIS_LONG:
        k, vals, rows, cols, _ = fill_sparse_laso(S_rows, S_cols, vec_nnz, state, dt)
        S = np.zeros((S_rows, S_cols), dtype=dt)
        np.add.at(S, (rows, cols), vals)
        nxt = laso_next_state(S_rows, S_cols, vec_nnz, state)
    else:
        tall = S_rows > S_cols
        wr, wc = (S_cols, S_rows) if tall else (S_rows, S_cols)
        k, vals, rows, cols, _ = fill_sparse(wr, wc, vec_nnz, state, dt)
        S = np.zeros((wr, wc), dtype=dt)
        np.add.at(S, (rows, cols), vals)
        if tall:
            S = S.T
        nxt = saso_next_state(wr, wc, vec_nnz, state)
    rs, cs = ((m, d) if opS else (d, m)) if left else ((d, n) if opS else (n, d))
    if S_rows < rs + ro or S_cols < cs + co:
        raise ValueError("sketch_general: submatrix of S out of range (randblas_require)")
    Sub = S[ro:ro + rs, co:co + cs]
    opSm = Sub.T if opS else Sub
    ra, ca = (n, m) if opA else (m, n)
    A2 = _mat_view(A_flat, ra, ca, lda, layout)
    opAm = A2.T if opA else A2
    R = (opSm @ opAm) if left else (opAm @ opSm)
    out = np.array(B_flat, copy=True)
    Bv = _mat_view(out, *((d, n) if left else (m, d)), ldb, layout)
    Bv[:, :] = dt.type(alpha) * R + (dt.type(beta) * Bv if beta != 0 else 0)
    return out, nxt


def col_swap(A, idx):
    """util::col_swap = lapack::lapmt(forward) (rl_util.hh:151-165): new column i = old column idx[i]-1."""
    return _F(A[:, np.asarray(idx, dtype=np.int64) - 1])


class CQRRPT:
    """RandLAPACK::CQRRPT(timing, eps); fields nnz (=2), rank, qrcp ('geqp3' default | 'bqrrp' | 'hqrrp', rl_cqrrpt.hh:230-247) and the
    HQRRP fields nb_alg, oversampling, panel_pivoting, use_cholqr (:134-137, defaults :60-63)."""

    def __init__(self, eps, nnz=2):
        self.eps, self.nnz, self.rank, self.qrcp, self.orthogonalization = eps, nnz, None, "geqp3", False
        self.nb_alg, self.oversampling, self.panel_pivoting, self.use_cholqr = 64, 10, 1, 0

    def call(self, A, d_factor, state: RNGState, R=None):
        """-> (rc, Q (m x n, first rank columns meaningful), R (n x n), J (1-based), next state)."""
        A = _F(np.array(A, copy=True))
        m, n = A.shape
        dt = A.dtype
        eps_m = np.finfo(dt).eps
        R = np.zeros((n, n), dtype=dt, order="F") if R is None else _F(np.array(R, copy=True))
        J = np.zeros(n, dtype=np.int64)
        k = n
        d = int(dt.type(d_factor) * dt.type(n))                                     # :197
        eps_initial = dt.type(2) * dt.type(np.power(np.float64(eps_m), 0.95))       # :199
        A_hat, state = sketch_sparse_left(d, m, self.nnz, d, A, state)              # :214-221
        geqp3, trsm_, potrf = get_lapack_funcs(("geqp3",), (A_hat,))[0], get_blas_funcs(("trsm",), (A,))[0], \
            get_lapack_funcs(("potrf",), (A,))[0]
        if self.qrcp == "bqrrp":                                                    # :232-244
            ratio = 1.0 if n <= 2000 else (0.5 if n <= 8000 else 1.0 / 32.0)
            bq = BQRRP(int(dt.type(n) * dt.type(ratio)))
            _, A_hat, tau, jpvt, state = bq.call(A_hat, 1.0, state)
        elif self.qrcp == "hqrrp":                                                  # :230-231
            _, A_hat, tau, jpvt, state = hqrrp(A_hat, self.nb_alg, self.oversampling, self.panel_pivoting, self.use_cholqr, state)
        else:
            A_hat, jpvt, tau, _, info = geqp3(A_hat)                                # :247
        J[:] = jpvt
        if not A_hat[0, 0]:                                                         # :256
            return 0, A, R, J, state
        dg = np.abs(np.diag(A_hat)[:n])
        for i in range(n):                                                          # :267-272
            if dg[i] / dg[0] < eps_initial:
                k = i
                break
        self.rank = k
        new_rank = k
        R[:k, :k] = np.triu(A_hat[:k, :k]) + np.tril(R[:k, :k], -1)                 # lacpy(Upper) :284
        A = col_swap(A, J)                                                          # :291-292
        if np.any(np.diag(R)[:k] == 0):                                             # :300-305
            return 1, A, R, J, state
        A[:, :k] = trsm_(1.0, _F(R[:k, :k]), _F(A[:, :k]), side=1, lower=0)          # :306
        G = np.triu(_gemm(_F(A[:, :k]), _F(A[:, :k]), ta=True))                     # syrk(Upper) :309
        c, info = potrf(_F(G + np.tril(R[:k, :k], -1)), lower=0, clean=0)           # :311
        self.potrf_info = int(info)        # (diagnostic for the tests: whether the a-posteriori rank estimate below ran)
        R[:k, :k] = c
        if info:                                                                    # :311-336
            running_max = running_min = R[0, 0]
            cond_threshold = np.sqrt(dt.type(self.eps) / eps_m)
            for i in range(k):
                curr = abs(R[i, i])
                running_max, running_min = max(running_max, curr), min(running_min, curr)
                if running_min * cond_threshold < running_max and i > 1:
                    new_rank = i - 1
                    break
        self.rank = new_rank                                                        # :339
        A[:, :new_rank] = trsm_(1.0, _F(R[:new_rank, :new_rank]), _F(A[:, :new_rank]), side=1, lower=0)   # :342
        if not self.orthogonalization:
            R[:new_rank, :] = R[:new_rank, :] @ np.triu(A_hat[:n, :n])              # trmm :345
        elif new_rank != n:                                                         # complete the orthonormal set (:347-368)
            cols = n - new_rank
            G, _ = fill_dense(m, cols, state, dtype=dt)          # (:351 discards fill_dense's return value: the state does not advance)
            G = _F(G)
            Qr = _F(A[:, :new_rank])
            temp = _gemm(Qr, G, ta=True)
            G = _F(G - Qr @ temp)
            geqrf, orgqr = get_lapack_funcs(("geqrf", "orgqr"), (G,))
            qr, tau_o, _, _ = geqrf(G)
            A[:, new_rank:] = orgqr(qr, tau_o)[0]
        return 0, A, R, J, state


class CQRRT:
    """RandLAPACK::CQRRT(timing, eps) (rl_cqrrt.hh:39-297); fields nnz (= 2), orthogonalization (False), compute_Q (True)."""

    def __init__(self, eps=None, nnz=2):
        self.eps, self.nnz, self.orthogonalization, self.compute_Q = eps, nnz, False, True

    def call(self, A, d_factor, state: RNGState, R=None):
        """-> (rc, Q m x n, R n x n, next state)."""
        A = _F(np.array(A, copy=True))
        m, n = A.shape
        dt = A.dtype
        R = np.zeros((n, n), dtype=dt, order="F") if R is None else _F(np.array(R, copy=True))
        d = int(dt.type(d_factor) * dt.type(n))                                     # :135
        A_hat, state = sketch_sparse_left(d, m, self.nnz, d, A, state)              # :144-152
        geqrf, potrf = get_lapack_funcs(("geqrf", "potrf"), (A_hat,))
        (trsm_,) = get_blas_funcs(("trsm",), (A,))
        wq = geqrf(A_hat, lwork=-1)
        A_hat, tau = geqrf(A_hat, lwork=int(wq[-2][0]))[:2]                         # :160
        R[:n, :n] = np.triu(A_hat[:n, :n]) + np.tril(R[:n, :n], -1)                 # lacpy(Upper) :167
        if np.any(np.diag(R) == 0):                                                 # :173-177
            return 1, A, R, state
        A = trsm_(1.0, _F(R), A, side=1, lower=0)                                   # :178
        G = np.triu(_gemm(A, A, ta=True))                                           # syrk(Upper) :186
        c, info = potrf(_F(G + np.tril(R, -1)), lower=0, clean=0)                   # :194
        R[:, :] = c
        if info:
            return 1, A, R, state
        if self.compute_Q:
            A = trsm_(1.0, _F(R), A, side=1, lower=0)                               # :238
        if not self.orthogonalization:
            R[:, :] = np.triu(R) @ np.triu(A_hat[:n, :n]) + np.tril(R, -1)          # trmm :249
        return 0, A, R, state


# --------------------------------------------------------------------------------------------
# SYPS / SYRF / REVD2 (RandLAPACK/comps/rl_syps.hh, comps/rl_syrf.hh, drivers/rl_revd2.hh)
# --------------------------------------------------------------------------------------------
def _symm(uplo, A, B):
    """ExplicitSymLinOp::operator() (linops/rl_sym_linops.hh:76-99): blas::symm(Left, uplo) - only the `uplo` triangle of A is read."""
    (symm,) = get_blas_funcs(("symm",), (A, B))
    return _F(symm(1.0, A, B, side=0, lower=int(uplo in ("L", "l", 1))))


class SYPS:
    """rl_syps.hh:21-143 - skop = fill_dense(DenseDist(m, k)); p passes of skop <- A skop with geqrf + ungqr every q passes."""

    def __init__(self, p, q, verbose=False, cond_check=False):
        self.passes_over_data, self.passes_per_stab = p, q

    def call(self, uplo, A, k, state: RNGState):
        """-> (0, skop m x k, next state)."""
        m = A.shape[0]
        p, q = self.passes_over_data, self.passes_per_stab
        mat, state = fill_dense(m, k, state, dtype=A.dtype)                          # :74-75
        skop = _F(mat)                          # m >= k: the natural layout of DenseDist(m, k) is column-major (dense_skops.hh:184-196)
        work = np.zeros((m, k), dtype=A.dtype, order="F")                            # :80
        geqrf, orgqr = get_lapack_funcs(("geqrf", "orgqr"), (A,))
        bufs = {"skop": skop, "work": work}
        out, inn = "work", "skop"
        p_done = 0
        while p - p_done > 0:
            bufs[out] = _symm(uplo, A, bufs[inn])                                    # :86
            p_done += 1
            if p_done % q == 0:                                                      # :88-93
                qr, tau, _, info = geqrf(bufs[out])
                if info:
                    raise RuntimeError("GEQRF failed.")
                bufs[out] = _F(orgqr(qr, tau)[0])
            out, inn = ("skop", "work") if p_done % 2 == 1 else ("work", "skop")     # :95-96
        if p % 2 == 1:
            bufs["skop"] = bufs["work"].copy(order="F")                              # :99-100
        return 0, bufs["skop"], state


class SYRF:
    """rl_syrf.hh:21-118 - Q = orth(A * syps(A)); a failing orthogonaliser raises as the reference throws."""

    def __init__(self, syps, orth, verbose=False, cond_check=False):
        self.syps, self.orth = syps, orth

    def call(self, uplo, A, k, state: RNGState):
        _, omega, state = self.syps.call(uplo, A, k, state)                          # :83
        Q = _symm(uplo, A, omega)                                                    # :86
        rc, Q = self.orth.call(Q)                                                    # :94
        if rc:
            raise RuntimeError("Orthogonalization failed.")
        return 0, Q, state


def power_error_est(uplo, A, k, p, g, V, eigvals):
    """rl_revd2.hh:20-71 (g: the m-vector the reference keeps in the first column of vector_buf)."""
    err = A.dtype.type(0)
    g = g.copy()
    for _ in range(p):
        g = g / np.linalg.norm(g)                                                    # :34-36
        t1 = V.T @ g                                                                 # :40
        Mat = V * eigvals[None, :k]                                                  # :44-48
        t2 = Mat @ t1                                                                # :52
        t3 = _symm(uplo, A, _F(g.reshape(-1, 1))).ravel()                            # :55
        w = t3 - t2                                                                  # :60
        err = g @ w                                                                  # :62
        g = w                                                                        # :64
    return err


class REVD2:
    """rl_revd2.hh:75-246."""

    def __init__(self, syrf, error_est_power_iters, verbose=False):
        self.syrf, self.error_est_p, self.err, self.k_history = syrf, error_est_power_iters, None, []

    def call(self, uplo, A, k, tol, state: RNGState):
        """-> (0, k, V m x k, eigvals k, next state)."""
        if not (k > 0 and tol >= 0):
            raise ValueError("randlapack_require failed")                            # :131-134
        A = _F(A)
        m = A.shape[0]
        dt = A.dtype
        est_state = RNGState(key=_key_incr(state.key, 1), counter=state.counter)     # :152-153
        (trsm,) = get_blas_funcs(("trsm",), (A,))
        (potrf,) = get_lapack_funcs(("potrf",), (A,))
        self.k_history = []
        while True:
            self.k_history.append(k)
            _, Omega, state = self.syrf.call(uplo, A, k, state)                      # :166
            Y = _symm(uplo, A, Omega)                                                # :169
            nu = dt.type(np.finfo(dt).eps) * dt.type(np.linalg.norm(Y, "fro"))       # :171
            R = nu * (Omega.T @ Omega)                                               # syrk + mirror (:177-179)
            R = _F(Omega.T @ Y + R)                                                  # :181
            c, info = potrf(R, lower=0, clean=1)                                     # :185-187 (get_U)
            if info:
                raise RuntimeError("Cholesky decomposition failed.")
            B = trsm(1.0, _F(c), Y, side=1, lower=0, trans_a=0, diag=0)              # :190
            V, S, _ = scipy.linalg.svd(B, full_matrices=False, lapack_driver="gesdd")   # :195
            V = _F(V)
            eig = (S * S).astype(dt)                                                 # :198-207
            r = int(np.sum(eig > nu))
            for i in range(r):                                                       # :211-212
                if not (eig[i] - nu < 0):
                    eig[i] -= nu
            V[:, r:] = 0                                                             # :214
            g, est_state = fill_dense(m, 1, est_state, dtype=dt)                     # :219-221
            err = power_error_est(uplo, A, k, self.error_est_p, np.asarray(g).ravel(order="A"), V, eig)   # :223
            self.err = err
            if err <= 5 * max(tol, nu) or k == m:                                    # :225-231
                break
            k = m if 2 * k > m else 2 * k
        return 0, k, V, eig, state


def _key_incr(key, n):
    v = (key[0] | (key[1] << 32)) + n
    return (v & 0xFFFFFFFF, (v >> 32) & 0xFFFFFFFF)


# --------------------------------------------------------------------------------------------
# BQRRP (RandLAPACK/drivers/rl_bqrrp.hh:154-665)
# --------------------------------------------------------------------------------------------
def orhr_col(Q):
    """LAPACK dorhr_col with a single T block (NB >= n), restated: Householder reconstruction of an m x n matrix with
    orthonormal columns.  -> (V unit-lower-trapezoidal m x n, T n x n upper, D (+-1)).
    Steps (LAPACK 3.9 dorhr_col.f / dlaorhr_col_getrfnp.f): modified LU without pivoting Q1 - S = L1 U1 with
    S = diag(D), D_i = -sign of the i-th pivot; V2 = Q2 U1^-1; T = (-U1 S) V1^-T.  Used at rl_bqrrp.hh:466."""
    from scipy.linalg import solve_triangular
    Q = np.array(Q, dtype=Q.dtype, order="F", copy=True)
    m, n = Q.shape
    D = np.zeros(n, dtype=Q.dtype)
    for i in range(n):
        D[i] = -np.copysign(1.0, Q[i, i])
        Q[i, i] -= D[i]
        Q[i + 1:n, i] *= Q.dtype.type(1.0) / Q[i, i]
        Q[i + 1:n, i + 1:n] -= np.outer(Q[i + 1:n, i], Q[i, i + 1:n])
    U1 = np.triu(Q[:n, :n])
    V1 = np.tril(Q[:n, :n], -1) + np.eye(n, dtype=Q.dtype)
    if m > n:
        Q[n:, :] = solve_triangular(U1, Q[n:, :].T, trans="T", lower=False).T
    V = Q.copy()
    V[:n, :n] = V1
    Tm = U1 * np.where(D == 1.0, -1.0, 1.0)[None, :]            # column j scaled by -D_j
    Tm = solve_triangular(V1, Tm.T, lower=True, unit_diagonal=True).T     # X V1^T = Tm
    return _F(V), _F(np.triu(Tm)), D


class BQRRP:
    """RandLAPACK::BQRRP(timing, b_sz): fields block_size, tol (= eps), qrcp_wide ('luqr' | 'geqp3'), qr_tall ('geqrf' | 'cholqr'), rank.
    apply_trans_q = ormqr (applied here through the compact-WY form, which is what ormqr computes)."""

    def __init__(self, b_sz, qrcp_wide="luqr", qr_tall="geqrf"):
        self.block_size, self.qrcp_wide, self.qr_tall, self.rank, self.tol = b_sz, qrcp_wide, qr_tall, None, None

    def call(self, A, d_factor, state: RNGState):
        """-> (rc, A_out [GEQP3 format], tau, J (1-based), next state)."""
        A = _F(np.array(A, copy=True))
        m, n = A.shape
        dt = A.dtype
        eps = np.finfo(dt).eps
        tol = eps if self.tol is None else dt.type(self.tol)                          # this->tol (:141), ctor default eps (:71)
        tau = np.zeros(n, dtype=dt)
        J = np.zeros(n, dtype=np.int64)
        rows, cols, curr, b_sz = m, n, 0, self.block_size
        maxiter = int(np.ceil(dt.type(min(m, n)) / dt.type(b_sz)))
        b_const = b_sz
        d = int(dt.type(d_factor) * dt.type(b_sz))
        sd = d
        geqp3, geqrf, getrf, potrf, ormqr = get_lapack_funcs(("geqp3", "geqrf", "getrf", "potrf", "ormqr"), (A,))
        (trsm_,) = get_blas_funcs(("trsm",), (A,))
        # sketch (:309-312): the row-major d x m fill_dense buffer is READ as ColMajor with ld = d
        S, state = fill_dense(d, m, state, dt)                   # natural layout of a wide Long-axis operator = row-major
        S_used = np.ascontiguousarray(S).reshape(-1).reshape((d, m), order="F")
        A_sk = _F(S_used @ A)
        sk0 = 0                                                  # column offset of the live sketch inside A_sk

        def ormqr_apply(V, tau_, C):
            _, lw, _ = ormqr("L", "T", V, tau_, C, lwork=-1)
            out, _, info = ormqr("L", "T", V, tau_, C, lwork=int(lw[0]))
            return out

        for it in range(maxiter):
            b_sz = min(b_sz, min(m, n) - curr)
            block_rank = b_sz
            Ask = A_sk[:sd, sk0:sk0 + cols]
            if self.qrcp_wide == "geqp3":                        # :335-336
                qr, jp, tw, _, _ = geqp3(_F(Ask))
                A_sk[:sd, sk0:sk0 + cols] = qr
                Jb = np.asarray(jp, dtype=np.int64)
            else:                                                # :337-357
                lu, piv, _ = getrf(_F(Ask.T))
                Jb = np.arange(1, cols + 1, dtype=np.int64)
                for i in range(min(sd, cols)):
                    Jb[piv[i]], Jb[i] = Jb[i], Jb[piv[i]]
                qr, tw, _, _ = geqrf(_F(Ask[:, Jb - 1]))
                A_sk[:sd, sk0:sk0 + cols] = qr
            A[:, curr:] = A[:, curr:][:, Jb - 1]                  # :365
            block_zero = not np.any(np.abs(A[curr:, curr]) > eps)   # :372-379
            if it == 0:
                J[:cols] = Jb
            else:
                J[curr:curr + cols] = J[curr:curr + cols][Jb - 1]
            if block_zero:
                self.rank = curr
                return 0, A, tau, J, state
            R_sk = A_sk[:, sk0:]
            for i in range(b_sz):                                # :421-427
                if abs(R_sk[i, i]) / abs(R_sk[0, 0]) < tol:
                    block_rank = i
                    break
            P = A[curr:, curr:curr + b_sz]                       # the panel (view)
            W1 = A[curr:, curr + b_sz:]
            if self.qr_tall == "cholqr":                         # :441-497
                br = block_rank
                X = trsm_(1.0, _F(np.triu(R_sk[:br, :br])), _F(P[:, :br]), side=1, lower=0)
                G = np.triu(_gemm(X, X, ta=True))
                c, info = potrf(_F(G), lower=0, clean=1)
                X = trsm_(1.0, _F(c), X, side=1, lower=0)
                V, Tm, D = orhr_col(X)
                Rfull = np.triu(c) * D[:, None]
                tau[curr:curr + br] = np.diag(Tm)
                Rpad = np.zeros((br, b_sz), dtype=dt)
                Rpad[:, :br] = Rfull
                R11 = Rpad @ np.triu(R_sk[:b_sz, :b_sz])         # trmm :486
                k_refl, Vk, tk = br, V, np.diag(Tm).copy()
                P[:, :br] = np.tril(V, -1)
                P[:br, :b_sz] = np.triu(R11) + np.tril(P[:br, :b_sz], -1)
            else:                                                # geqrf :498-510
                qr, tq, _, _ = geqrf(_F(P))
                P[:, :] = qr
                tau[curr:curr + len(tq)] = tq
                k_refl, Vk, tk = block_rank, qr, tq
            if k_refl > 0 and cols - b_sz > 0:                   # :541-562
                m_apply = block_rank if block_rank != b_const else rows
                Vq = _F(np.tril(Vk[:m_apply, :k_refl], -1) + np.eye(m_apply, k_refl, dtype=dt))
                W1[:m_apply, :] = ormqr_apply(Vq, np.asarray(tk[:k_refl], dtype=dt), _F(W1[:m_apply, :]))
            curr += b_sz
            if curr >= min(m, n) or block_rank != b_const:       # :583-598
                self.rank = curr
                return 0, A, tau, J, state
            R11 = np.triu(A[curr - b_sz:curr, curr - b_sz:curr])
            R12 = A[curr - b_sz:curr, curr:]
            R_sk[:b_sz, :b_sz] = np.triu(R_sk[:b_sz, :b_sz])     # get_U :605
            R_sk[:b_sz, :b_sz] = trsm_(1.0, _F(R11), _F(R_sk[:b_sz, :b_sz]), side=1, lower=0)   # :606
            R_sk[:b_sz, b_sz:cols] -= R_sk[:b_sz, :b_sz] @ R12   # :610
            sd = min(sd, cols)
            if sd - b_sz > 0:                                    # :617-618
                blk = R_sk[b_sz:sd, b_sz:sd]
                blk[:, :] = np.triu(blk)
            sk0 += b_sz
            rows -= b_sz
            cols -= b_sz
        self.rank = curr
        return 0, A, tau, J, state


# --------------------------------------------------------------------------------------------
# HQRRP (RandLAPACK/drivers/rl_hqrrp.hh): Householder QR with randomized pivoting
# --------------------------------------------------------------------------------------------
def _qrp_unb(A, p, t, num_stages, pivoting, B=None, C=None, build_T=False):
    """NoFLA_QRPmod_WY_unb_var4's unblocked loop (rl_hqrrp.hh:556-775): Householder QR of A (in place, a Fortran-ordered view) with
    optional column pivoting by downdated partial column norms; B and C (views with A's column count) are pivoted along; p / t are
    views of the pivot and tau vectors.  Returns T (larft, Forward / Columnwise, num_stages x num_stages) when build_T."""
    m, n = A.shape
    dt = A.dtype
    mn = min(m, n)
    if num_stages < 0:
        num_stages = mn
    larfg, = get_lapack_funcs(("larfg",), (A,))
    tol3z = np.sqrt(np.finfo(np.float64).eps / 2)                  # sqrt(dlamch('E')) for every T (:376-378)
    if pivoting:
        d = np.array([np.linalg.norm(A[:, j]) for j in range(n)], dtype=dt)    # NoFLA_QRP_compute_norms (:336-357)
        e = d.copy()
    for j in range(num_stages):
        if pivoting:
            jm = int(np.argmax(d[j:]))                             # blas::iamax: first maximum (:662)
            if jm != 0:                                            # NoFLA_QRP_pivot_G_B_C (:414-461)
                a, b = j, j + jm
                A[:, [a, b]] = A[:, [b, a]]
                if B is not None:
                    B[:, [a, b]] = B[:, [b, a]]
                if C is not None:
                    C[:, [a, b]] = C[:, [b, a]]
                p[a], p[b] = p[b], p[a]
                d[b], e[b] = d[a], e[a]                            # norms of column 0 are COPIED to column j_max_col
        # larfg on (alpha11, a21) (:683-688)
        if m - j - 1 > 0:
            alpha, x, tau_j = larfg(m - j, A[j, j], A[j + 1:, j].copy())
            A[j, j], A[j + 1:, j], t[j] = alpha, x, tau_j
        else:
            t[j] = 0                                               # larfg with n = 1: tau = 0
        # | a12t; A22 | = H | a12t; A22 | (:700-709)
        if n - j - 1 > 0 and t[j] != 0:
            v = np.concatenate(([dt.type(1)], A[j + 1:, j]))
            blk = A[j:, j + 1:]
            w = blk.T @ v
            blk -= t[j] * np.outer(v, w)
        if pivoting and n - j - 1 > 0:                             # NoFLA_QRP_downdate_partial_norms (:360-411)
            for c in range(j + 1, n):
                if d[c] != 0:
                    temp = abs(A[j, c]) / d[c]
                    temp = max(0.0, (1.0 + temp) * (1 - temp))
                    temp5 = d[c] / e[c]
                    temp2 = temp * temp5 * temp5
                    if temp2 <= tol3z:
                        d[c] = np.linalg.norm(A[j + 1:, c]) if m - j - 1 > 0 else 0
                        e[c] = d[c]
                    else:
                        d[c] = d[c] * np.sqrt(temp)
    if build_T:
        return _larft(A, t, num_stages)
    return None


def _larft(V, tau, k):
    """lapack::larft(Forward, Columnwise) on the unit-lower-trapezoidal reflectors stored below the diagonal of V (m x >= k)."""
    dt = V.dtype
    m = V.shape[0]
    Vu = np.tril(V[:, :k], -1) + np.eye(m, k, dtype=dt)
    T = np.zeros((k, k), dtype=dt, order="F")
    for i in range(k):
        T[i, i] = tau[i]
        if i > 0 and tau[i] != 0:
            T[:i, i] = -tau[i] * (T[:i, :i] @ (Vu[:, :i].T @ Vu[:, i]))
    return T


def hqrrp(A, nb_alg, pp, panel_pivoting, qr_type, state: RNGState):
    """RandLAPACK::hqrrp(m, n, A, lda, jpvt, tau, nb_alg, pp, panel_pivoting, qr_type, state, timing) (rl_hqrrp.hh:811-1196).
    -> (rc, A_out [GEQP3 format], tau (n entries, first min(m, n) meaningful), J (1-based), next state).
    qr_type (the panel QR when panel_pivoting == 0): 0 unblocked Householder, 1 geqrf (:464-502), 2 CholQR + orhr_col (:505-553)."""
    A = _F(np.array(A, copy=True))
    m, n = A.shape
    dt = A.dtype
    mn = min(m, n)
    tau = np.zeros(n, dtype=dt)
    J = np.zeros(n, dtype=np.int64)
    if mn == 0:                                                        # quick return (:886-888): J is not initialised
        return 0, A, tau, J, state
    m_Y = nb_alg + pp
    J[:] = np.arange(1, n + 1)                                         # std::iota (:919)
    # G = fill_dense(DenseDist(nb_alg + pp, m, Uniform)): the natural-layout buffer is read as ColMajor with ld = m_Y (:928-935)
    Gm, state = fill_dense(m_Y, m, state, dt, family=FAMILY_UNIFORM)
    G = _F(Gm.ravel(order="K").reshape((m_Y, m), order="F"))
    Y = _F(G @ A)
    geqrf, potrf = get_lapack_funcs(("geqrf", "potrf"), (A,))
    (trsm_,) = get_blas_funcs(("trsm",), (A,))
    for j in range(0, mn, nb_alg):
        b = min(nb_alg, n - j, m - j)
        last_iter = (j + nb_alg >= m) or (j + nb_alg >= n)
        if not last_iter:                                              # QRP of a copy of YR; AR and YR pivoted along (:1019-1046)
            V = _F(Y[:, j:].copy())
            _qrp_unb(V, J[j:], tau[j:], b, True, B=A[:, j:], C=Y[:, j:])
        AB1 = A[j:, j:j + b]
        if panel_pivoting:                                             # :1075-1080 with pivoting = 1
            Tm = _qrp_unb(AB1, J[j:j + b], tau[j:j + b], -1, True, B=A[:j, j:j + b], C=Y[:, j:j + b], build_T=True)
        elif qr_type == 2:                                             # CHOLQR_mod_WY (:505-553)
            R = np.triu(_gemm(_F(AB1), _F(AB1), ta=True))
            c, info = potrf(_F(R), lower=0, clean=1)
            if info:      # CHOLQR_mod_WY returns 1 and hqrrp, which does not look at it (:1075), goes on with an unfactored panel
                raise NotImplementedError("hqrrp: Cholesky failure inside a panel (the reference's output is unspecified here)")
            Q = trsm_(1.0, _F(c), _F(AB1), side=1, lower=0)
            Vh, Tm, D = orhr_col(Q)
            Rs = np.triu(c) * D[:, None]
            AB1[:, :] = np.tril(Vh, -1)
            AB1[:b, :b] += np.triu(Rs)
            tau[j:j + b] = np.diag(Tm)
        else:                                                          # geqrf (qr_type 1, :464-502) or the unblocked loop without pivoting (qr_type 0)
            if qr_type == 1:
                qr, tq, _, _ = geqrf(_F(AB1))
                AB1[:, :] = qr
                tau[j:j + len(tq)] = tq
                Tm = _larft(AB1, tau[j:j + b], min(AB1.shape))
            else:
                Tm = _qrp_unb(AB1, J[j:j + b], tau[j:j + b], -1, False, build_T=True)
        k = Tm.shape[0]
        if j + b < n:                                                  # [A12; A22] <- Q^T [A12; A22] (:1091-1100), larfb Left / Trans
            U = np.tril(AB1[:, :k], -1) + np.eye(m - j, k, dtype=dt)
            C2 = A[j:, j + b:]
            C2 -= U @ (Tm.T @ (U.T @ C2))
        if not last_iter:                                              # NoFLA_Downdate_Y (:206-296)
            U11 = np.tril(AB1[:b, :b], -1) + np.eye(b, dtype=dt)
            U21 = AB1[b:, :b]
            G1, G2 = G[:, j:j + b], G[:, j + b:]
            Bm = ((G1 @ U11 + G2 @ U21) @ Tm) @ U11.T
            Bm = G1 - Bm
            Y[:, j + b:] -= Bm @ A[j:j + b, j + b:]
            GR = G[:, j:]                                              # GR <- GR Q, larfb Right / NoTrans (:288-291)
            Ufull = np.vstack((U11, U21))
            GR -= ((GR @ Ufull) @ Tm) @ Ufull.T
    return 0, A, tau, J, state
