"""GPU parity, BQRRP through the C-ABI.

Tolerances (stated): return code, rank and RNG state exact; the pivot vector J bit-exact over the numerical rank against the golden
vectors of the real reference (columns past the numerical rank have |R_ii| ~ 1e-15 and order round-off noise — also between the
reference and its own restatement); diag(R), tau and the leading block of the GEQP3-formatted output to 1e-9 (fp64) / 2e-3 (fp32)
relative; and the reference's own acceptance test (test/drivers/test_bqrrp.cc:62-107): all three measures <= eps^0.75.
qr_tall = geqrf (the reference's default) and cholqr (BQRRP_GPU's choice) produce the same Householder representation up to
round-off for full-rank panels; both are exercised."""
import numpy as np
import pytest
import torch

import randlapack_b200 as rl
from _qrcases import G, bq_input, geqp3_format_invariants, numerical_rank
from oracle import rl_oracle as O

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.array(a.T, order="C", copy=True)).cuda().t()


def host(t):
    return np.asfortranarray(t.cpu().numpy())


def _run(ctx, A, c, st, qr_tall=None):
    alg = rl.BQRRP(False, c["b"])
    alg.qrcp_wide = c["qrcp_wide"]
    alg.qr_tall = c["qr_tall"] if qr_tall is None else qr_tall
    Ad = dev(A)
    s = rl.RNGState(st.key, st.counter)
    rc, tau, J = alg.call(ctx, Ad, c["d_factor"], s)
    return rc, alg.rank, host(Ad), tau.cpu().numpy(), J.cpu().numpy(), s


@pytest.mark.parametrize("i", range(int(G["bq_count"])))
def test_bqrrp_golden(ctx, i):
    A, st, c = bq_input(i)
    rc, rank, F, tau, J, s = _run(ctx, A, c, st)
    rc_ref, rank_ref = [int(x) for x in G[f"bq{i}_rc_rank"]]
    assert (rc, rank) == (rc_ref, rank_ref)
    assert list(s.words()) == list(G[f"bq{i}_state_out"])
    assert sorted(J.tolist()) == list(range(1, c["n"] + 1))
    kn = min(numerical_rank(F), rank_ref)
    assert np.array_equal(J[:kn], G[f"bq{i}_J"][:kn]), "pivot vector differs from the reference's"
    tol = 1e-9 if c["dtype"] == np.float64 else 2e-3
    dref = G[f"bq{i}_Rdiag"]
    assert np.abs(np.diag(F)[:kn] - dref[:kn]).max() <= tol * np.abs(dref).max()
    assert np.abs(tau[:kn] - G[f"bq{i}_tau"][:kn]).max() <= tol * 10
    h = min(48, kn)
    assert np.abs(F[:h, :h] - G[f"bq{i}_Fhead"][:h, :h]).max() <= tol * 10 * np.abs(dref).max()
    e = geqp3_format_invariants(A, F, tau, J, rank if kn == rank_ref else kn)
    atol = np.finfo(c["dtype"]).eps ** 0.75
    assert e[2] <= atol and (kn < rank_ref or max(e) <= atol), e


@pytest.mark.parametrize("i", [0, 3])
def test_bqrrp_cholqr_equals_geqrf_representation(ctx, i):
    """The CholQR + Householder-reconstruction panel gives the reflectors geqrf gives (full-rank panels)."""
    A, st, c = bq_input(i)
    out0 = _run(ctx, A, c, st, qr_tall=rl.QRTALL_GEQRF)
    out1 = _run(ctx, A, c, st, qr_tall=rl.QRTALL_CHOLQR)
    assert out0[0] == out1[0] and out0[1] == out1[1]
    # all panels but the last: identical pivots and representation (in the last panel of the wide case the candidate columns are
    # nearly dependent on the 512 already factored ones and round-off decides)
    k = min(c["m"], c["n"]) - c["b"] if c["m"] < c["n"] else c["n"]
    assert np.array_equal(out0[4][:k], out1[4][:k])
    sc = np.abs(np.diag(out0[2])).max()
    assert np.abs(out0[2][:, :k] - out1[2][:, :k]).max() <= 1e-9 * sc and np.abs(out0[3][:k] - out1[3][:k]).max() <= 1e-9


@pytest.mark.parametrize("shape", [(1, 1, 1), (7, 3, 2), (40, 40, 8), (100, 37, 16), (37, 100, 16), (300, 64, 64), (300, 64, 100)])
def test_bqrrp_edge_shapes_vs_oracle(ctx, shape):
    m, n, b = shape
    rng = np.random.default_rng(m * 31 + n)
    A = np.asfortranarray(rng.standard_normal((m, n)) * (1.0 + 0.3 * np.arange(n))[None, :] ** -1.0)
    for qw, qt in ((0, 0), (1, 1), (0, 1)):
        c = dict(b=b, qrcp_wide=qw, qr_tall=qt, d_factor=1.0, n=n, dtype=np.float64)
        if b > m:      # DenseDist(d, m) must be wide
            continue
        rc, rank, F, tau, J, s = _run(ctx, A, c, O.RNGState(0))
        o = O.BQRRP(b, "geqp3" if qw else "luqr", "cholqr" if qt else "geqrf")
        rc2, F2, tau2, J2, st2 = o.call(A, 1.0, O.RNGState(0))
        # (wide inputs: once the m rows are exhausted the LU of the rank-deficient last sketch pivots on round-off, so only the
        #  first min(m, n) pivots are defined)
        assert (rc, rank) == (rc2, o.rank) and np.array_equal(J[:min(m, n)], J2[:min(m, n)]), (shape, qw, qt)
        assert sorted(J.tolist()) == list(range(1, n + 1))
        assert list(s.words()) == list(st2.words())
        r = min(rank, m, n)
        assert np.abs(np.triu(F)[:r, :r] - np.triu(F2)[:r, :r]).max() <= 1e-9 * np.abs(F2).max()
        assert np.abs(tau - tau2).max() <= 1e-9
        e = geqp3_format_invariants(A, F, tau, J, r)
        assert max(e) <= np.finfo(np.float64).eps ** 0.75, (shape, qw, qt, e)


def test_bqrrp_zero_matrix_and_host_call(ctx):
    Z = rl.to_f(torch.zeros((200, 50), dtype=torch.float64, device="cuda"))
    alg = rl.BQRRP(False, 16)
    rc, tau, J = alg.call(ctx, Z, 1.0, rl.RNGState(0))
    assert rc == 0 and alg.rank == 0 and float(Z.abs().max()) == 0.0       # test_bqrrp.cc:128-132
    A, st, c = bq_input(0)
    Ah = torch.from_numpy(np.array(A.T, order="C", copy=True)).t()
    alg = rl.BQRRP(False, c["b"])
    s = rl.RNGState(st.key, st.counter)
    rc, tau, J = alg.call_host(ctx, Ah, c["d_factor"], s)
    assert rc == 0 and alg.rank == int(G["bq0_rc_rank"][1]) and np.array_equal(J.numpy(), G["bq0_J"])
    with pytest.raises(rl.Error):
        alg.call(ctx, Z, 0.5, rl.RNGState(0))                                # randlapack_require(d_factor >= 1)


def test_bqrrp_large_property(ctx):
    """8192 x 4096 fp64, b = 256 (config C4's block size), BQRRP_GPU's configuration: J a permutation, |diag R| non-increasing up to
    the sketch's distortion, and ||A[:,J]^T A[:,J] - R^T R|| small (R^T R = A^T A for any QR, no Q needed)."""
    m, n, b = 8192, 4096, 256
    g = torch.Generator(device="cuda").manual_seed(7)
    A = rl.to_f(torch.randn((m, n), dtype=torch.float64, device="cuda", generator=g))
    A *= (1.0 + torch.arange(n, device="cuda", dtype=torch.float64))[None, :] ** -0.25
    A0 = A.clone()
    alg = rl.BQRRP(False, b)
    alg.qr_tall = rl.QRTALL_CHOLQR
    rc, tau, J = alg.call(ctx, A, 1.0, rl.RNGState(3))
    assert rc == 0 and alg.rank == n
    assert sorted(J.cpu().tolist()) == list(range(1, n + 1))
    R = torch.triu(A[:n])
    AP = A0[:, J - 1]
    G1 = AP.t() @ AP
    G2 = R.t() @ R
    assert float((G1 - G2).norm() / G1.norm()) <= 1e-12
    dg = R.diagonal().abs()
    blocks = dg.view(-1, b).max(dim=1).values
    assert bool((blocks[1:] <= blocks[:-1] * 1.5).all())


@pytest.mark.parametrize("qr_tall", ["geqrf", "cholqr"])
def test_bqrrp_engine_trailing_update_vs_oracle(ctx, qr_tall):
    """the trailing update only runs on the int8 digit-slice engine when rows - k >= 8192 (bqrrp.cu), which no r1 test
    reached.  20000 x 768, b = 256: the first two panels update 19744 / 19488 rows on tcgen05.  Against the ORACLE (numpy/LAPACK restatement
    of rl_bqrrp.hh) on the same input and state: code, rank, state and pivots exact; R and tau to 1e-9; the reference's eps^0.75 test."""
    m, n, b = 20000, 768, 256
    A, st = O.gen_poly_mat(m, n, n, 1.0e4, 2.0, O.RNGState(0))
    c = dict(b=b, qrcp_wide=0, qr_tall=1 if qr_tall == "cholqr" else 0, d_factor=1.0, n=n, dtype=np.float64)
    rc, rank, F, tau, J, s = _run(ctx, A, c, st)
    o = O.BQRRP(b, "luqr", qr_tall)
    rc2, F2, tau2, J2, st2 = o.call(A, 1.0, O.RNGState(st.key, st.counter))
    assert (rc, rank) == (rc2, o.rank)
    assert list(s.words()) == list(st2.words())
    assert np.array_equal(J, J2), "pivot vector differs from the oracle's"
    sc = np.abs(np.diag(F2)).max()
    assert np.abs(np.triu(F)[:n, :n] - np.triu(F2)[:n, :n]).max() <= 1e-9 * sc
    assert np.abs(tau - tau2).max() <= 1e-9
    e = geqp3_format_invariants(A, F, tau, J, rank)
    assert max(e) <= np.finfo(np.float64).eps ** 0.75, e


@pytest.mark.parametrize("shape,b,d_factor,qr_tall", [((2000, 1500, 1500), 128, 1.0, "geqrf"), ((2000, 1500, 1500), 128, 1.0, "cholqr"),
                                                      ((1000, 1000, 100), 64, 2.0, "geqrf"), ((20000, 512, 512), 256, 1.0, "cholqr")])
def test_bqrrp_gpu_interface_same_sketch_as_cpu(ctx, shape, b, d_factor, qr_tall):
    """BQRRP_GPU_alg::call(m, n, A, lda, A_sk, d, tau, J) (rl_bqrrp_gpu.hh:27-43, 122-133): device pointers, the sketch is an INPUT.
    The reference's own GPU-vs-CPU recipe (test/drivers/test_bqrrp_gpu.cu:91-103, 224-249): S = fill_dense(DenseDist(d, m)) read as a
    column-major d x m matrix, A_sk = S A formed on the host, handed to the device driver; the CPU driver (here: the oracle's restatement of
    rl_bqrrp.hh) regenerates the same sketch from the same state.  ||J_gpu - J_cpu|| = 0, tau to eps^0.75, R to eps^0.60."""
    m, n, k = shape
    A, st = O.gen_poly_mat(m, n, k, 2025.0, 2.0, O.RNGState(0))
    d = int(d_factor * b)
    S, _ = O.fill_dense(d, m, O.RNGState(st.key, st.counter))
    # the reference hands the natural-layout (row-major) buffer to gemm as ColMajor with ld = d (rl_bqrrp.hh:309-312, test_bqrrp_gpu.cu:99)
    S = np.ascontiguousarray(S).reshape(-1).reshape((d, m), order="F")
    A_sk = np.asfortranarray(S @ A)
    o = O.BQRRP(b, "luqr", qr_tall)
    rc2, F2, tau2, J2, _ = o.call(A, d_factor, O.RNGState(st.key, st.counter))
    alg = rl.BQRRP(False, b)
    alg.qr_tall = 1 if qr_tall == "cholqr" else 0
    Ad, Skd = dev(A), dev(A_sk)
    rc, tau, J = alg.call_sk(ctx, Ad, Skd)
    F, tau, J = host(Ad), tau.cpu().numpy(), J.cpu().numpy()
    assert (rc, alg.rank) == (rc2, o.rank)
    kn = min(numerical_rank(F2), o.rank)
    assert np.array_equal(J[:kn], J2[:kn]), "pivot vector differs from the CPU driver's on the same sketch"
    eps = np.finfo(np.float64).eps
    assert np.linalg.norm(tau[:kn] - tau2[:kn]) <= eps ** 0.75 * 10
    sc = np.abs(np.diag(F2)).max()
    # columns past the numerical rank are ordered by round-off noise (on both sides): compare the part of R they do not permute
    nc = n if kn == min(m, n) else kn
    assert np.linalg.norm(np.triu(F[:kn, :nc]) - np.triu(F2[:kn, :nc])) <= eps ** 0.60 * sc
    e = geqp3_format_invariants(A, F, tau, J, alg.rank if kn == o.rank else kn)
    assert e[2] <= eps ** 0.75 and (kn < o.rank or max(e) <= eps ** 0.75), e


def test_bqrrp_tol_field_vs_oracle(ctx):
    """BQRRP's public `tol` (rl_bqrrp.hh:141, rank cut at :422): a rank-70 matrix plus 1e-7 noise factors to full rank with the default
    tol = eps and is cut at 96 columns with tol = 1e-3 - the same rank, pivots and R as the oracle (itself pinned to the compiled
    reference in tests/test_oracle_qr.py::test_bqrrp_tol_field)."""
    rng = np.random.default_rng(3)
    A = np.asfortranarray(rng.standard_normal((600, 70)) @ rng.standard_normal((70, 200)) + 1e-7 * rng.standard_normal((600, 200)))
    for tol, want in ((1e-3, 96), (None, 200)):
        alg = rl.BQRRP(False, 32)
        alg.tol = tol
        Ad = dev(A)
        s = rl.RNGState(4)
        rc, tau, J = alg.call(ctx, Ad, 1.0, s)
        o = O.BQRRP(32)
        o.tol = tol
        rc2, F2, tau2, J2, st2 = o.call(A, 1.0, O.RNGState(4))
        assert (rc, alg.rank) == (rc2, o.rank) and alg.rank == want
        assert list(s.words()) == list(st2.words())
        r = min(alg.rank, 70)      # beyond the numerical rank the pivots order noise
        assert np.array_equal(J.cpu().numpy()[:r], J2[:r])
        F = host(Ad)
        assert np.abs(np.triu(F)[:r, :] - np.triu(F2)[:r, :]).max() <= 1e-9 * np.abs(np.diag(F2)).max()
    ctx.check(ctx._lib.rlb200_set_bqrrp_tol(ctx._h, 0.0))
