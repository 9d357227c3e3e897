"""GPU parity, CQRRPT and its building blocks through the C-ABI.

Tolerances (stated): return code, rank, RNG state and the pivot vector J are exact (bit-exact integers) against the golden vectors
of the real reference and against the oracle on matrices whose pivot gaps dominate round-off (for rank-deficient inputs only the
first `rank` pivots are meaningful — the rest order round-off noise, also between the reference and its own restatement).
R and Q: fp64 1e-9 relative / fp32 2e-3 (CholQR of a sketch-preconditioned matrix; the sketch is summed in another order), and
the reference's own acceptance test (test/drivers/test_cqrrpt.cc:98-104): all three error measures <= eps^0.75."""
import numpy as np
import pytest
import torch

import randlapack_b200 as rl
from _qrcases import G, cq_input, qr_invariants
from oracle import rl_oracle as O

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a.T)).cuda().t()


def host(t):
    return np.asfortranarray(t.cpu().numpy())


@pytest.mark.parametrize("i", range(int(G["cq_count"])))
def test_cqrrpt_golden(ctx, i):
    A, st, c = cq_input(i)
    alg = rl.CQRRPT(False, c["eps"])
    alg.nnz = c["nnz"]
    Ad = dev(A)
    s = rl.RNGState(st.key, st.counter)
    rc, R, J = alg.call(ctx, Ad, c["d_factor"], s)
    rc_ref, rank_ref = [int(x) for x in G[f"cq{i}_rc_rank"]]
    assert (rc, alg.rank) == (rc_ref, rank_ref)
    assert list(s.words()) == list(G[f"cq{i}_state_out"])
    Q, R, J = host(Ad), host(R), J.cpu().numpy()
    r = rank_ref
    assert sorted(J.tolist()) == list(range(1, c["n"] + 1))
    assert np.array_equal(J[:r], G[f"cq{i}_J"][:r]), "pivot vector differs from the reference's"
    if np.array_equal(J, G[f"cq{i}_J"]):
        tol = 1e-9 if c["dtype"] == np.float64 else 2e-3
        assert np.abs(np.diag(R)[:r] - G[f"cq{i}_Rdiag"][:r]).max() <= tol * np.abs(G[f"cq{i}_Rdiag"]).max()
        h = min(r, 32)
        assert np.abs(R[:h] - G[f"cq{i}_Rhead"]).max() <= tol * np.abs(G[f"cq{i}_Rhead"]).max()
        assert np.abs(Q[:32, :r] - G[f"cq{i}_Qhead"]).max() <= tol * 10
    e = qr_invariants(A, Q, R, J, r)
    assert max(e) <= np.finfo(c["dtype"]).eps ** 0.75, e


def test_cqrrpt_host_call_and_zero_matrix(ctx):
    A, st, c = cq_input(2)
    alg = rl.CQRRPT(False, c["eps"])
    Ah = torch.from_numpy(np.array(A.T, order="C", copy=True)).t()   # (A.T of an F-ordered array is already C-contiguous: force a copy)
    s = rl.RNGState(st.key, st.counter)
    rc, R, J = alg.call_host(ctx, Ah, c["d_factor"], s)
    assert rc == 0 and alg.rank == int(G["cq2_rc_rank"][1]) and np.array_equal(J.numpy(), G["cq2_J"])
    e = qr_invariants(A, Ah.numpy(), R.numpy(), J.numpy(), alg.rank)
    assert max(e) <= np.finfo(np.float64).eps ** 0.75
    # all-zero input: the reference returns 0 right after the QRCP (rl_cqrrpt.hh:256-261), A untouched
    Z = rl.to_f(torch.zeros((500, 20), dtype=torch.float64, device="cuda"))
    rc, R, J = alg.call(ctx, Z, 2.0, rl.RNGState(0))
    assert rc == 0 and float(Z.abs().max()) == 0.0
    # argument validation = randlapack_require (rl_cqrrpt.hh:161-168)
    with pytest.raises(rl.Error):
        alg.call(ctx, Z, 0.5, rl.RNGState(0))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("shape", [(64, 40), (300, 300), (40, 64), (1500, 700), (1, 1), (5, 1)])
def test_qrcp_small_vs_lapack(ctx, dtype, shape):
    """geqp3 of the sketch: pivots bit-exact against LAPACK's on columns with graded norms; R equal to round-off; and geqrf."""
    from scipy.linalg import lapack
    d, n = shape
    npdt = np.float64 if dtype == torch.float64 else np.float32
    rng = np.random.default_rng(d * 7 + n)
    A = rng.standard_normal((d, n)) * (1.0 + 0.37 * rng.permutation(n))[None, :] ** -1.0
    A = np.asfortranarray(A.astype(npdt))
    Ad = dev(A)
    J, tau = rl.qr_small(ctx, Ad, pivot=True)
    f = lapack.dgeqp3 if npdt == np.float64 else lapack.sgeqp3
    qr, jpvt, tau_ref, _, info = f(A)
    assert np.array_equal(J.cpu().numpy(), jpvt)
    tol = 1e-11 if npdt == np.float64 else 1e-3
    Rd, Rr = np.triu(host(Ad))[:min(d, n)], np.triu(qr)[:min(d, n)]
    assert np.abs(Rd - Rr).max() <= tol * np.abs(Rr).max()
    assert np.abs(tau.cpu().numpy() - tau_ref).max() <= tol * 10
    assert np.abs(np.tril(host(Ad), -1) - np.tril(qr, -1)).max() <= tol * 100
    # unpivoted
    Ad = dev(A)
    _, tau = rl.qr_small(ctx, Ad, pivot=False)
    g = lapack.dgeqrf if npdt == np.float64 else lapack.sgeqrf
    qr, tau_ref, _, info = g(A)
    assert np.abs(host(Ad) - qr).max() <= tol * 100 * np.abs(qr).max()
    assert np.abs(tau.cpu().numpy() - tau_ref).max() <= tol * 10


def test_qrcp_ties_and_zero_columns(ctx):
    """iamax semantics: first maximum wins; zero columns stay at the end with tau = 0."""
    from scipy.linalg import lapack
    A = np.zeros((8, 6), order="F")
    A[:, 1] = 1.0
    A[:, 3] = 1.0           # exact tie with column 1
    A[0, 4] = 0.5
    Ad = dev(A)
    J, tau = rl.qr_small(ctx, Ad, pivot=True)
    qr, jpvt, tau_ref, _, info = lapack.dgeqp3(A)
    assert np.array_equal(J.cpu().numpy(), jpvt)
    assert np.allclose(tau.cpu().numpy(), tau_ref, atol=1e-14)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_col_swap(ctx, dtype):
    rng = np.random.default_rng(3)
    for (m, n) in [(1, 1), (7, 5), (1000, 33), (4096, 64), (1001, 17)]:
        A = rl.to_f(torch.randn((m, n), dtype=dtype, device="cuda"))
        ref = A.clone()
        idx = rng.permutation(n) + 1
        rl.col_swap(ctx, A, idx)
        assert torch.equal(A, ref[:, torch.from_numpy(idx - 1).cuda()])
    # property at scale: a permutation followed by its inverse is the identity (bit-exact), on a matrix larger than L2
    m, n = 1 << 20, 64
    A = rl.to_f(torch.randn((m, n), dtype=dtype, device="cuda"))
    ref = A.clone()
    idx = rng.permutation(n) + 1
    inv = np.argsort(idx - 1) + 1
    rl.col_swap(ctx, A, idx)
    assert not torch.equal(A, ref)
    rl.col_swap(ctx, A, inv)
    assert torch.equal(A, ref)


def test_cqrrpt_large_property(ctx):
    """A size the oracle does not run at (2^20 x 512 fp32, n > 256 exercises the blocked Cholesky / right-solve):
    ||Q'Q - I||, ||A[:,J] - QR|| on a row sample, J a permutation, R upper triangular."""
    m, n = 1 << 20, 512
    g = torch.Generator(device="cuda").manual_seed(5)
    A = rl.to_f(torch.randn((m, n), dtype=torch.float32, device="cuda", generator=g))
    A *= (1.0 + torch.arange(n, device="cuda", dtype=torch.float32))[None, :] ** -0.5
    A0 = A.clone()
    alg = rl.CQRRPT(False, None)
    alg.nnz = 1
    rc, R, J = alg.call(ctx, A, 2.0, rl.RNGState(1))
    assert rc == 0 and alg.rank == n
    assert sorted(J.cpu().tolist()) == list(range(1, n + 1))
    assert float(torch.tril(R, -1).abs().max()) == 0.0
    QtQ = rl.gemm(ctx, True, False, 1.0, A, A)
    assert float((QtQ - torch.eye(n, device="cuda")).norm()) / n ** 0.5 <= 1e-4
    rows = torch.randint(0, m, (4096,), device="cuda")
    E = A0[rows][:, J - 1].double() - A[rows].double() @ R.double()
    assert float(E.norm() / A0[rows].double().norm()) <= 1e-4


@pytest.mark.parametrize("dtype,m,n", [(torch.float64, 40000, 300), (torch.float32, 40000, 300), (torch.float32, 1 << 18, 512)])
def test_cqrrpt_i8_engine(ctx, dtype, m, n):
    """CQRRPT with its O(m n^2) work (A R^-1 as one in-place product with the explicit inverse, Gram matrix) on the tcgen05 int8
    digit-slice engine: same return code, rank and pivots (bit-exact) as the fp64-pipe path, R to the path's tolerance, and the
    reference's acceptance test (test_cqrrpt.cc:98-104) on the result."""
    g = torch.Generator(device="cuda").manual_seed(11)
    A0 = rl.to_f(torch.randn((m, n), dtype=dtype, device="cuda", generator=g))
    A0 *= (1.0 + torch.arange(n, device="cuda", dtype=dtype))[None, :] ** -0.75
    outs = []
    for engine in ("dmma", "i8"):
        ctx.set_fp64_engine(engine)
        try:
            A = A0.clone()
            alg = rl.CQRRPT(False, None)
            alg.nnz = 2
            rc, R, J = alg.call(ctx, A, 1.5, rl.RNGState(7))
        finally:
            ctx.set_fp64_engine("i8")
        outs.append((rc, alg.rank, J.cpu().numpy(), R, A))
    (rc0, rk0, J0, R0, Q0), (rc1, rk1, J1, R1, Q1) = outs
    assert (rc0, rk0) == (rc1, rk1) == (0, n)
    assert np.array_equal(J0, J1)
    tol = 1e-8 if dtype == torch.float64 else 2e-3
    assert float((R1 - R0).abs().max() / R0.abs().max()) <= tol
    eps = np.finfo(np.float64 if dtype == torch.float64 else np.float32).eps
    QtQ = rl.gemm(ctx, True, False, 1.0, Q1, Q1)
    assert float((QtQ - torch.eye(n, device="cuda", dtype=dtype)).norm()) / n ** 0.5 <= eps ** 0.75
    rows = torch.randint(0, m, (4096,), device="cuda")
    Jt = torch.from_numpy(J1 - 1).cuda()
    E = A0[rows][:, Jt].double() - Q1[rows].double() @ R1.double()
    assert float(E.norm() / A0[rows].double().norm()) <= eps ** 0.75


@pytest.mark.parametrize("cond", [1.0e2, 1.0e10])
def test_cqrrpt_engine_vs_oracle_blocked_sizes(ctx, cond):
    """CQRRPT on the int8 engine against the ORACLE (not the repo's own DMMA path) at n > 300, where the blocked
    Cholesky and the in-place product with the explicit inverse R^-1 run (drivers.cu: tall_right_solve / tall_gram_upper), with a
    well- and an ill-conditioned sketch factor (cond(A) = 1e10: R^-1 has entries ~1e10, the stress case for the explicit inverse).
    Pivots, rank, code, state exact; R to 1e-8 relative; the reference's three eps^0.75 measures (test_cqrrpt.cc:98-104) computed in numpy."""
    m, n = 40000, 640
    A, st = O.gen_poly_mat(m, n, n, cond, 2.0, O.RNGState(0))
    alg = rl.CQRRPT(False, None)
    alg.nnz = 2
    Ad = dev(A)
    s = rl.RNGState(st.key, st.counter)
    rc, R, J = alg.call(ctx, Ad, 1.25, s)
    o = O.CQRRPT(np.finfo(np.float64).eps ** 0.85, nnz=2)
    rc2, Q2, R2, J2, st2 = o.call(A, 1.25, O.RNGState(st.key, st.counter))
    assert (rc, alg.rank) == (rc2, o.rank)
    assert list(s.words()) == list(st2.words())
    Q, R, J = host(Ad), host(R), J.cpu().numpy()
    r = alg.rank
    assert np.array_equal(J[:r], np.asarray(J2)[:r]), "pivot vector differs from the oracle's"
    assert np.abs(R[:r] - R2[:r]).max() <= 1e-8 * np.abs(R2).max()
    e = qr_invariants(A, Q, R, J, r)
    assert max(e) <= np.finfo(np.float64).eps ** 0.75, e


# ---------------------------------------------------------------------------------------------------------------------------
# CQRRT (rl_cqrrt.hh:91-297): unpivoted sketched Cholesky QR on the same kernels
# ---------------------------------------------------------------------------------------------------------------------------
from _qrcases import GT, ct_input, check_cqrrt_against_golden  # noqa: E402


@pytest.mark.parametrize("i", range(int(GT["ct_count"])))
def test_cqrrt_golden(ctx, i):
    A, st, c = ct_input(i)
    alg = rl.CQRRT(False, None)
    alg.nnz, alg.orthogonalization, alg.compute_Q = c["nnz"], c["orth"], c["compute_Q"]
    Ad = dev(A)
    s = rl.RNGState(st.key, st.counter)
    rc, R = alg.call(ctx, Ad, c["d_factor"], s)
    check_cqrrt_against_golden(i, c, A, rc, host(Ad), host(R), s.words())


def test_cqrrt_engine_vs_oracle_and_host_call(ctx):
    """Tall enough for the tall products to run on the tcgen05 digit-slice engine (m >= 16384), n > 256 (blocked Cholesky / right-solve),
    against the oracle on the same input and state; then the host-pointer entry (the reference's calling convention)."""
    m, n = 40000, 320
    A, st = O.gen_poly_mat(m, n, n, 1.0e3, 2.0, O.RNGState(0))
    o = O.CQRRT(None, 2)
    rc_o, Q_o, R_o, st_o = o.call(A, 1.5, O.RNGState(st.key, st.counter))
    alg = rl.CQRRT(False, None)
    Ad = dev(A)
    s = rl.RNGState(st.key, st.counter)
    rc, R = alg.call(ctx, Ad, 1.5, s)
    assert rc == rc_o == 0 and list(s.words()) == list(st_o.words())
    Q, R = host(Ad), host(R)
    sc = np.abs(np.diag(R_o)).max()
    assert np.abs(np.triu(R) - np.triu(R_o)).max() <= 1e-9 * sc
    assert np.abs(Q - Q_o).max() <= 1e-9
    e = qr_invariants(A, Q, np.triu(R), np.arange(1, n + 1), n)
    assert max(e) <= np.finfo(np.float64).eps ** 0.75, e
    Ah = torch.from_numpy(np.ascontiguousarray(A.T)).t()
    s2 = rl.RNGState(st.key, st.counter)
    rc2, R2 = alg.call_host(ctx, Ah, 1.5, s2)
    assert rc2 == 0 and np.abs(np.triu(R2.numpy()) - np.triu(R)).max() <= 1e-12 * sc


def test_cqrrt_zero_column_returns_1(ctx):
    A, st = O.gen_poly_mat(500, 20, 20, 10.0, 2.0, O.RNGState(0))
    A[:, 7] = 0
    rc, _ = rl.CQRRT(False, None).call(ctx, dev(A), 2.0, rl.RNGState(st.key, st.counter))
    assert rc == 1


def test_phase_times_vectors(ctx):
    """The reference's public `times` vectors (rl_cqrrpt.hh:371-384: 8 entries; rl_cqrrt.hh:279-282: 10; rl_bqrrp.hh:582-584: 10): same
    length and order, microseconds, the last entry the total and the entries before it summing to it."""
    A, st = O.gen_poly_mat(20000, 128, 128, 100.0, 2.0, O.RNGState(0))
    ctx.phase_timing(True)
    try:
        rl.CQRRPT(True, None).call(ctx, dev(A), 1.5, rl.RNGState(st.key, st.counter))
        t = ctx.phase_times()
        assert len(t) == 8 and t[-1] > 0 and sum(t[:-1]) == t[-1] and all(x >= 0 for x in t[:6])
        rl.CQRRT(True, None).call(ctx, dev(A), 1.5, rl.RNGState(st.key, st.counter))
        t = ctx.phase_times()
        assert len(t) == 10 and t[2] == 0 and t[5] == 0 and t[-1] > 0 and sum(t[:-1]) == t[-1]
        alg = rl.BQRRP(True, 64)
        alg.call(ctx, dev(A[:4000]), 1.0, rl.RNGState(st.key, st.counter))
        t = ctx.phase_times()
        assert len(t) == 10 and t[-1] > 0 and sum(t[:-1]) == t[-1]
    finally:
        ctx.phase_timing(False)


@pytest.mark.parametrize("shape", [(3000, 200, 1e3, 1.5), (20000, 300, 1e6, 2.0)])
def test_cqrrpt_qrcp_bqrrp_vs_oracle(ctx, shape):
    """CQRRPT's `qrcp` field = bqrrp (rl_cqrrpt.hh:41, 232-244: the QRCP of the sketch by BQRRP(false, n * ratio), which draws its own sketch from
    the state): rank, pivots and the advanced state exact, R 1e-9, the reference's eps^0.75 measures - against the oracle, itself pinned to
    the compiled reference (tests/test_oracle_qr.py::test_cqrrpt_qrcp_bqrrp)."""
    m, n, cond, df = shape
    A, st = O.gen_poly_mat(m, n, n, cond, 2.0, O.RNGState(0))
    alg = rl.CQRRPT(False, None)
    alg.qrcp = "bqrrp"
    Ad = dev(A)
    s = rl.RNGState(st.key, st.counter)
    rc, R, J = alg.call(ctx, Ad, df, s)
    o = O.CQRRPT(float(np.finfo(np.float64).eps) ** 0.85, 2)
    o.qrcp = "bqrrp"
    rc2, Q2, R2, J2, st2 = o.call(A, df, O.RNGState(st.key, st.counter))
    assert (rc, alg.rank) == (rc2, o.rank)
    assert list(s.words()) == list(st2.words())
    J, R, Q = J.cpu().numpy(), host(R), host(Ad)
    assert np.array_equal(J, J2)
    k = alg.rank
    assert np.abs(np.triu(R[:k]) - np.triu(R2[:k])).max() <= 1e-9 * np.abs(R2).max()
    e = qr_invariants(A, Q, R, J, k)
    assert max(e) <= np.finfo(np.float64).eps ** 0.75, e
    alg.qrcp = "lu"                                   # not one of the reference's three choices (rl_cqrrpt.hh:39-43)
    with pytest.raises(rl.Error):
        alg.call(ctx, dev(A), df, rl.RNGState(0))
    ctx.check(ctx._lib.rlb200_set_cqrrpt_qrcp(ctx._h, 0))


@pytest.mark.parametrize("shape", [(3000, 200, 120, 1e3, 1.5), (2000, 120, 120, 10.0, 1.25), (20000, 150, 60, 1e2, 2.0)])
def test_cqrrpt_orthogonalization_mode_vs_oracle(ctx, shape):
    """CQRRPT's `orthogonalization` field (rl_cqrrpt.hh:139-142, 343-368) against the oracle (pinned to the compiled reference in
    tests/test_oracle_qr.py): rank, leading pivots and state exact; R (the Cholesky factor, preconditioning not undone) 1e-9; all n columns
    of Q orthonormal, the first `rank` ones equal to the oracle's up to sign and the completed ones orthogonal to them."""
    m, n, rk, cond, df = shape
    A, st = O.gen_poly_mat(m, n, rk, cond, 2.0, O.RNGState(0))
    alg = rl.CQRRPT(False, None)
    alg.orthogonalization = True
    Ad = dev(A)
    s = rl.RNGState(st.key, st.counter)
    rc, R, J = alg.call(ctx, Ad, df, s)
    o = O.CQRRPT(float(np.finfo(np.float64).eps) ** 0.85, 2)
    o.orthogonalization = True
    rc2, Q2, R2, J2, st2 = o.call(A, df, O.RNGState(st.key, st.counter))
    k = o.rank
    assert (rc, alg.rank) == (rc2, k) and list(s.words()) == list(st2.words())
    J, R, Q = J.cpu().numpy(), host(R), host(Ad)
    assert np.array_equal(J[:k], J2[:k])
    assert np.abs(np.triu(R[:k, :k]) - np.triu(R2[:k, :k])).max() <= 1e-9 * np.abs(R2).max()
    assert np.linalg.norm(Q.T @ Q - np.eye(n)) <= 1e-11
    assert np.abs(np.abs(Q[:, :k]) - np.abs(Q2[:, :k])).max() <= 1e-8
    if k < n:      # same Gaussian block, same projection: the completed columns span the oracle's completion
        P1, P2 = Q[:, k:] @ Q[:, k:].T, Q2[:, k:] @ Q2[:, k:].T
        v = np.random.default_rng(0).standard_normal((m, 3))
        assert np.abs(P1 @ v - P2 @ v).max() <= 1e-9 * np.linalg.norm(v)      # (the leading Q itself agrees to ~1e-9: CholQR of a sketch-preconditioned matrix)
    ctx.check(ctx._lib.rlb200_set_cqrrpt_orthogonalization(ctx._h, 0))
