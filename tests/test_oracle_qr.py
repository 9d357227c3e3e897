"""CPU: the oracle's CQRRPT restatement against the golden vectors from the real reference (tests/golden/qr_vectors.npz) and,
when oracle/_ref exists, against the compiled reference directly.  Pivots, rank, return code and RNG state are exact;
R and Q to round-off (the sketch is summed in a different order than the reference's OpenMP-simd loop)."""
import numpy as np
import pytest

import _ref
from _qrcases import G, cq_input, qr_invariants
from oracle import rl_oracle as O


@pytest.mark.parametrize("i", range(int(G["cq_count"])))
def test_cqrrpt_golden(i):
    A, st, c = cq_input(i)
    alg = O.CQRRPT(c["eps"], c["nnz"])
    rc, Q, R, J, st2 = alg.call(A, c["d_factor"], st)
    rc_ref, rank_ref = [int(x) for x in G[f"cq{i}_rc_rank"]]
    assert (rc, alg.rank) == (rc_ref, rank_ref)
    assert list(st2.words()) == list(G[f"cq{i}_state_out"])
    r = rank_ref
    full = r == c["n"]
    assert np.array_equal(J[:r], G[f"cq{i}_J"][:r]) or not full     # beyond the numerical rank the pivots are round-off noise
    if full or np.array_equal(J, G[f"cq{i}_J"]):
        tol = 1e-9 if c["dtype"] == np.float64 else 2e-3
        assert np.abs(np.diag(R)[:r] - G[f"cq{i}_Rdiag"][:r]).max() <= tol * np.abs(G[f"cq{i}_Rdiag"]).max()
        assert np.abs(Q[:32, :r] - G[f"cq{i}_Qhead"]).max() <= tol * 10
    e = qr_invariants(A, Q, R, J, r)
    atol = np.finfo(c["dtype"]).eps ** 0.75
    assert max(e) <= atol, e


@pytest.mark.skipif(_ref.ref_lib() is None, reason="compiled reference not present")
def test_cqrrpt_vs_compiled_reference():
    L = _ref.ref_lib()
    for (m, n, k, cond, dt, nnz, df) in [(10, 5, 5, 2, np.float64, 2, 2.0), (1500, 80, 80, 1e3, np.float64, 3, 1.5), (900, 64, 64, 50, np.float32, 4, 1.25)]:
        A, st = _ref.ref_mat_gen(L, 0, m, n, k, cond, 2.0, [0] * 6, dt)
        eps = float(np.finfo(dt).eps) ** 0.85
        rc, rank, Q, R, J, st2 = _ref.ref_cqrrpt(L, A, df, st, eps, nnz)
        o = O.CQRRPT(eps, nnz)
        rc2, Q2, R2, J2, st3 = o.call(A, df, O.RNGState((st[4], st[5]), st[:4]))
        assert (rc, rank) == (rc2, o.rank) and np.array_equal(J, J2) and st2 == list(st3.words())
        tol = 1e-11 if dt == np.float64 else 1e-4
        assert np.abs(R - R2).max() <= tol * np.abs(R).max() and np.abs(Q[:, :rank] - Q2[:, :rank]).max() <= tol * 10


from _qrcases import bq_input, geqp3_format_invariants, numerical_rank  # noqa: E402


@pytest.mark.parametrize("i", range(int(G["bq_count"])))
def test_bqrrp_golden(i):
    """Oracle BQRRP vs the real reference's output: return code, rank, RNG state exact; pivots exact over the numerical rank;
    R diagonal / tau / leading block to round-off; the reference's acceptance invariants (test_bqrrp.cc:104-107)."""
    A, st, c = bq_input(i)
    alg = O.BQRRP(c["b"], "geqp3" if c["qrcp_wide"] else "luqr", "cholqr" if c["qr_tall"] else "geqrf")
    rc, F, tau, J, st2 = alg.call(A, c["d_factor"], st)
    rc_ref, rank_ref = [int(x) for x in G[f"bq{i}_rc_rank"]]
    assert (rc, alg.rank) == (rc_ref, rank_ref)
    assert list(st2.words()) == list(G[f"bq{i}_state_out"])
    kn = min(numerical_rank(F), rank_ref)
    assert np.array_equal(J[:kn], G[f"bq{i}_J"][:kn])
    tol = 1e-9 if c["dtype"] == np.float64 else 2e-3
    dref = G[f"bq{i}_Rdiag"]
    assert np.abs(np.diag(F)[:kn] - dref[:kn]).max() <= tol * np.abs(dref).max()
    assert np.abs(tau[:kn] - G[f"bq{i}_tau"][:kn]).max() <= tol * 10
    h = min(48, kn)
    assert np.abs(F[:h, :h] - G[f"bq{i}_Fhead"][:h, :h]).max() <= tol * 10 * np.abs(dref).max()
    e = geqp3_format_invariants(A, F, tau, J, min(alg.rank, kn) if kn < rank_ref else alg.rank)
    atol = np.finfo(c["dtype"]).eps ** 0.75
    assert e[2] <= atol and (kn < rank_ref or max(e) <= atol), e


# ---------------------------------------------------------------------------------------------------------------------------
# CQRRT (rl_cqrrt.hh:91-297): the oracle's restatement against the goldens of the compiled reference
# ---------------------------------------------------------------------------------------------------------------------------
from _qrcases import GT, ct_input, check_cqrrt_against_golden  # noqa: E402


@pytest.mark.parametrize("i", range(int(GT["ct_count"])))
def test_cqrrt_golden(i):
    A, st, c = ct_input(i)
    alg = O.CQRRT(None, c["nnz"])
    alg.orthogonalization, alg.compute_Q = c["orth"], c["compute_Q"]
    rc, Q, R, st2 = alg.call(A, c["d_factor"], st)
    check_cqrrt_against_golden(i, c, A, rc, Q, R, st2.words())


def test_cqrrt_zero_column_returns_1():
    """rl_cqrrt.hh:173-177: a zero diagonal entry of the sketch's R (here: a zero column) -> return 1."""
    A, st = O.gen_poly_mat(500, 20, 20, 10.0, 2.0, O.RNGState(0))
    A[:, 7] = 0
    rc, *_ = O.CQRRT(None, 2).call(A, 2.0, st)
    assert rc == 1


@pytest.mark.skipif(_ref.ref_lib() is None, reason="compiled reference not present")
def test_bqrrp_tol_field():
    """The oracle's `tol` (rl_bqrrp.hh:141, :422) against the compiled reference with the same public field set."""
    L = _ref.ref_lib()
    rng = np.random.default_rng(3)
    A = np.asfortranarray(rng.standard_normal((600, 70)) @ rng.standard_normal((70, 200)) + 1e-7 * rng.standard_normal((600, 200)))
    for tol, want in ((1e-3, 96), (None, 200)):
        st = O.RNGState(4)
        rc, rank, F, tau, J, st2 = _ref.ref_bqrrp(L, A, 1.0, 32, list(st.words()), tol=tol)
        alg = O.BQRRP(32)
        alg.tol = tol
        rc2, F2, tau2, J2, st3 = alg.call(A, 1.0, st)
        assert (rc, rank) == (rc2, alg.rank) and rank == want and list(st3.words()) == st2
        r = min(rank, 70)
        assert np.array_equal(J[:r], J2[:r])
        assert np.abs(np.triu(F)[:r] - np.triu(F2)[:r]).max() <= 1e-10 * np.abs(np.diag(F)).max()


@pytest.mark.skipif(_ref.ref_lib() is None, reason="compiled reference not present")
@pytest.mark.parametrize("shape", [(3000, 200, 1e3, 1.5), (2000, 120, 10.0, 1.25)])
def test_cqrrpt_qrcp_bqrrp(shape):
    """CQRRPT with qrcp = bqrrp (rl_cqrrpt.hh:41, 232-244): the oracle against the compiled reference with the same field set."""
    L = _ref.ref_lib()
    m, n, cond, df = shape
    A, st = O.gen_poly_mat(m, n, n, cond, 2.0, O.RNGState(0))
    rc, rank, Q, Rm, J, st2 = _ref.ref_cqrrpt(L, A, df, list(st.words()), None, 2, qrcp=1)
    alg = O.CQRRPT(float(np.finfo(np.float64).eps) ** 0.85, 2)
    alg.qrcp = "bqrrp"
    rc2, Q2, R2, J2, st3 = alg.call(A, df, st)
    assert (rc, rank) == (rc2, alg.rank) and list(st3.words()) == st2 and np.array_equal(J, J2)
    assert np.abs(np.triu(Rm) - np.triu(R2)).max() <= 1e-12 * np.abs(Rm).max()


@pytest.mark.skipif(_ref.ref_lib() is None, reason="compiled reference not present")
@pytest.mark.parametrize("shape", [(3000, 200, 120, 1e3, 1.5), (2000, 120, 120, 10.0, 1.25), (4000, 150, 60, 1e2, 2.0)])
def test_cqrrpt_orthogonalization_mode(shape):
    """CQRRPT::orthogonalization (rl_cqrrpt.hh:139-142, 343-368): R keeps the Cholesky factor, rank-deficient inputs get their Q completed to n
    orthonormal columns from a Gaussian block drawn at the (not advanced) state.  Oracle vs the compiled reference."""
    L = _ref.ref_lib()
    m, n, rk, cond, df = shape
    A, st = O.gen_poly_mat(m, n, rk, cond, 2.0, O.RNGState(0))
    rc, rank, Q, Rm, J, st2 = _ref.ref_cqrrpt(L, A, df, list(st.words()), None, 2, qrcp=0, orthogonalization=True)
    alg = O.CQRRPT(float(np.finfo(np.float64).eps) ** 0.85, 2)
    alg.orthogonalization = True
    rc2, Q2, R2, J2, st3 = alg.call(A, df, st)
    assert (rc, rank) == (rc2, alg.rank) and list(st3.words()) == st2 and np.array_equal(J[:rank], J2[:rank])
    assert np.abs(np.triu(Rm[:rank, :rank]) - np.triu(R2[:rank, :rank])).max() <= 1e-11 * np.abs(Rm).max()
    assert np.abs(np.abs(Q) - np.abs(Q2)).max() <= 1e-10
    assert np.linalg.norm(Q2.T @ Q2 - np.eye(n)) <= 1e-12


from _qrcases import GH, check_hqrrp_against_golden, hq_input  # noqa: E402


@pytest.mark.parametrize("i", range(int(GH["hq_count"])))
def test_hqrrp_golden(i):
    """The restatement of RandLAPACK::hqrrp (rl_hqrrp.hh:811-1196) against golden vectors of the compiled reference
    (tests/golden/make_golden_hqrrp.py): code, RNG state and pivots exact, R / tau to round-off, test_hqrrp.cc's acceptance measures."""
    A, st, c = hq_input(i)
    rc, F, tau, J, st2 = O.hqrrp(A, c["nb_alg"], c["pp"], c["panel_pivoting"], c["qr_type"], st)
    check_hqrrp_against_golden(i, c, A, rc, F, tau, J, st2.words())


def test_hqrrp_quick_return_and_tall_operator():
    """min(m, n) = 0 returns before the pivot vector is initialised (rl_hqrrp.hh:886-888); nb_alg + pp > m makes DenseDist(nb_alg + pp, m)
    a TALL operator whose natural layout is column-major."""
    rc, F, tau, J, st = O.hqrrp(np.zeros((0, 5)), 4, 2, 1, 0, O.RNGState(3))
    assert rc == 0 and not J.any() and list(st.words()) == list(O.RNGState(3).words())
    L = _ref.ref_lib()
    if L is None:
        return
    A, s0 = _ref.ref_mat_gen(L, 0, 40, 30, 30, 10.0, 2.0, [0] * 6, np.float64)
    rc, F, tau, J, s1 = _ref.ref_hqrrp(L, A, 64, 10, 1, 0, s0)
    rc2, F2, tau2, J2, st2 = O.hqrrp(A, 64, 10, 1, 0, O.RNGState.from_words(s0))
    assert rc == rc2 and np.array_equal(J, J2) and list(st2.words()) == s1
    assert np.abs(np.triu(F[:30]) - np.triu(F2[:30])).max() <= 1e-12


@pytest.mark.skipif(_ref.ref_lib() is None, reason="compiled reference not present")
@pytest.mark.parametrize("shape", [(3000, 200, 1e3, 1.5, 64, 10), (2000, 120, 10.0, 1.25, 32, 8)])
def test_cqrrpt_qrcp_hqrrp(shape):
    """CQRRPT with qrcp = hqrrp (rl_cqrrpt.hh:41, 230-231): the oracle against the compiled reference with the same field set
    (test/drivers/test_cqrrpt.cc:184-304 runs this option on 10000 x 200)."""
    L = _ref.ref_lib()
    m, n, cond, df, nb, ov = shape
    A, st = O.gen_poly_mat(m, n, n, cond, 2.0, O.RNGState(0))
    alg = O.CQRRPT(float(np.finfo(np.float64).eps) ** 0.85, 2)
    alg.qrcp = "hqrrp"
    if (nb, ov) == (64, 10):          # the compiled reference's constructor defaults
        rc, rank, Q, Rm, J, st2 = _ref.ref_cqrrpt(L, A, df, list(st.words()), None, 2, qrcp=2)
    else:
        alg.nb_alg, alg.oversampling = nb, ov
    rc2, Q2, R2, J2, st3 = alg.call(A, df, st)
    if (nb, ov) == (64, 10):
        assert (rc, rank) == (rc2, alg.rank) and list(st3.words()) == st2 and np.array_equal(J, J2)
        assert np.abs(np.triu(Rm) - np.triu(R2)).max() <= 1e-12 * np.abs(Rm).max()
    e = qr_invariants(A, Q2, R2, J2, alg.rank)
    assert alg.rank == n and max(e) <= np.finfo(np.float64).eps ** 0.75


@pytest.mark.skipif(_ref.ref_lib() is None, reason="compiled reference not present")
def test_cqrrpt_qrcp_hqrrp_reference_test_shape():
    """The reference's own CQRRPT-with-HQRRP case (test/drivers/test_cqrrpt.cc:184-304: 10000 x 200, rank 100, d_factor 2): the restatement
    against the compiled reference.  Rank and RNG state exact; the pivots are defined up to the last complete HQRRP block below the rank
    (nb_alg = 64: past it the sketch's columns are round-off), R's leading block to 1e-10."""
    L = _ref.ref_lib()
    m, n, rk = 10000, 200, 100
    A, st = O.gen_poly_mat(m, n, rk, 2.0, 2.0, O.RNGState(0))
    eps = float(np.finfo(np.float64).eps) ** 0.85
    rc, rank, Q, Rm, J, st2 = _ref.ref_cqrrpt(L, A, 2.0, list(st.words()), eps, 2, qrcp=2)
    alg = O.CQRRPT(eps, 2)
    alg.qrcp = "hqrrp"
    rc2, Q2, R2, J2, st3 = alg.call(A, 2.0, st)
    assert (rc, rank) == (rc2, alg.rank) == (0, rk) and list(st3.words()) == st2
    assert np.array_equal(J[:64], J2[:64])
    assert np.abs(np.triu(Rm[:64, :64]) - np.triu(R2[:64, :64])).max() <= 1e-10 * np.abs(Rm).max()
    e = qr_invariants(A, Q2, R2, J2, alg.rank)
    assert max(e) <= np.finfo(np.float64).eps ** 0.75, e
