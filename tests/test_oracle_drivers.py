"""The numpy restatement of RS/RF/QB/RSVD (oracle/rl_oracle.py) against the real reference:
golden fixtures generated from oracle/_ref (always) and the compiled reference itself (when present).
Also re-expresses the reference's own invariants for this path (test/comps/test_orth.cc:98,
test/comps/test_qb.cc:162-174)."""
import os

import numpy as np
import pytest

import _ref
from oracle import rl_oracle as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


def _inputs(i):
    m, n, k, p, q, b, stab = [int(x) for x in GOLD[f"rsvd{i}_args"]]
    cond, expo = GOLD[f"rsvd{i}_cond_expo"]
    A, st = O.gen_poly_mat(m, n, n if m > 10 else k, cond, expo, O.RNGState(0))
    return (m, n, k, p, q, b, stab), A, st


@pytest.mark.parametrize("i", range(int(GOLD["rsvd_count"])))
def test_rsvd_golden(i):
    (m, n, k, p, q, b, stab), A, st = _inputs(i)
    chk = GOLD[f"rsvd{i}_Achk"]
    assert np.allclose([A.sum(), np.abs(A).sum(), A[0, 0], A[-1, -1]], chk, rtol=1e-12, atol=1e-13)   # same input matrix
    assert list(st.counter) + list(st.key) == [int(x) for x in GOLD[f"rsvd{i}_state_in"]]
    *_, rsvd = O.make_stack(O.StackOpts(p, q, b, stab, 1, 1, False, False))
    rc, kk, U, S, V, st2 = rsvd.call(A, k, 0.0, st)
    assert [rc, kk] == [int(x) for x in GOLD[f"rsvd{i}_rc_k"]]
    assert list(st2.counter) + list(st2.key) == [int(x) for x in GOLD[f"rsvd{i}_state_out"]]
    # tolerance: fp64 round-off only (BLAS threading may reorder sums); singular vectors up to sign
    assert np.allclose(S, GOLD[f"rsvd{i}_S"], rtol=1e-10, atol=1e-13)
    resid = np.linalg.norm(A - (U * S) @ V.T) / np.linalg.norm(A)
    assert abs(resid - GOLD[f"rsvd{i}_resid"][0]) < 1e-10
    gaps = np.abs(np.diff(GOLD[f"rsvd{i}_S"]))
    if gaps.min() > 1e-6:   # non-degenerate spectrum: vectors are unique up to sign
        assert np.allclose(np.abs(V), np.abs(GOLD[f"rsvd{i}_V"]), atol=1e-7)
        assert np.allclose(np.abs(U[:64]), np.abs(GOLD[f"rsvd{i}_Uhead"]), atol=1e-7)


def test_rsvd_vs_compiled_reference():
    R = _ref.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    R.rlref_set_num_threads(1)
    for dt in (np.float64, np.float32):
        for (m, n, k, p, q, b, stab) in [(200, 60, 16, 0, 1, 16, 0), (200, 60, 16, 2, 1, 16, 0), (200, 60, 16, 3, 1, 8, 0),
                                         (300, 50, 20, 4, 2, 5, 1), (100, 100, 10, 5, 1, 10, 2), (10, 10, 5, 10, 1, 2, 0)]:
            A, st1 = _ref.ref_mat_gen(R, 0, m, n, n, 2025.0 if m > 10 else 2.0, 2.0, [0] * 6, dt)
            A2, st2 = O.gen_poly_mat(m, n, n, 2025.0 if m > 10 else 2.0, 2.0, O.RNGState(0), dtype=dt)
            assert np.array_equal(A, A2) and st1 == list(st2.counter) + list(st2.key)
            o = O.StackOpts(p, q, b, stab, 1, 1, False, False)
            rc, kk, U, S, V, s3 = _ref.ref_rsvd(R, A, k, 0.0, st1, o)
            *_, rsvd = O.make_stack(o)
            rc2, kk2, U2, S2, V2, s4 = rsvd.call(A2, k, 0.0, st2)
            assert (rc, kk) == (rc2, kk2) and s3 == list(s4.counter) + list(s4.key)
            tol = 1e-12 if dt == np.float64 else 1e-4
            assert np.allclose(S, S2, rtol=tol, atol=tol)
            assert np.allclose(np.abs(U), np.abs(U2), atol=1e3 * tol) and np.allclose(np.abs(V), np.abs(V2), atol=1e3 * tol)


def test_cholqrq_invariant():
    # test/comps/test_orth.cc:135-153: CholQRQ on a 1000x200 Gaussian sketch of a cond-2 matrix, applied twice,
    # ||Q'Q - I||_F <= eps^0.625 (:98)
    m, n, k = 1000, 200, 200
    A, st = O.gen_poly_mat(m, n, k, 2.0, 2.0, O.RNGState(0))
    Om, st = O.fill_dense(n, k, st)
    Y = np.asfortranarray(A @ Om)
    orth = O.CholQRQ()
    rc, Y = orth.call(Y)
    assert rc == 0
    rc, Y = orth.call(Y)
    assert rc == 0
    assert np.linalg.norm(Y.T @ Y - np.eye(k)) <= np.finfo(np.float64).eps ** 0.625


def test_cholqrq_failure_code():
    # rl_orth.hh:81-85: potrf failure => chol_fail = true, return 1
    Y = np.asfortranarray(np.ones((50, 4)))
    orth = O.CholQRQ()
    rc, _ = orth.call(Y)
    assert rc == 1 and orth.chol_fail


def test_qb_invariants():
    # test/comps/test_qb.cc:236-363: QB (CholQRQ x2) 100x100 rank-50; ||A - QB||, ||Q'Q - I|| <= eps^0.625 (:162-174)
    m = n = 100
    k = 50
    A, st = O.gen_poly_mat(m, n, k, 2025.0, 2.0, O.RNGState(0))
    _, _, qb, _ = O.make_stack(O.StackOpts(5, 1, 10, O.STAB_HQRQ, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ, False, True))
    rc, kk, Q, BT, st = qb.call(A, k, 10, 0.0, st)
    assert kk == k
    eps = np.finfo(np.float64).eps
    assert np.linalg.norm(A - Q @ BT.T) <= eps ** 0.625
    assert np.linalg.norm(Q.T @ Q - np.eye(k)) <= eps ** 0.625
