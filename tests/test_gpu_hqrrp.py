"""GPU parity, hqrrp (RandLAPACK/drivers/rl_hqrrp.hh:811-1196) through the C-ABI.

Tolerances (stated): return code and RNG state exact; the pivot vector J a permutation and bit-exact over the numerical rank against the golden
vectors of the compiled reference (tests/golden/hqrrp_vectors.npz) and against the restatement (oracle/rl_oracle.py::hqrrp, itself pinned to
the compiled reference); diag(R), tau and the leading block of the GEQP3-formatted output to 1e-9 (fp64) / 2e-3 (fp32) relative; the
reference's own acceptance test (test/drivers/test_hqrrp.cc:62-107): all three measures <= eps^0.75."""
import numpy as np
import pytest
import torch

import randlapack_b200 as rl
from _qrcases import GH, check_hqrrp_against_golden, geqp3_format_invariants, hq_input, qr_invariants
from oracle import rl_oracle as O

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.array(a.T, order="C", copy=True)).cuda().t()


def host(t):
    return np.asfortranarray(t.cpu().numpy())


def _run(ctx, A, c, st):
    Ad = dev(A)
    s = rl.RNGState(st.key, st.counter)
    rc, tau, J = rl.hqrrp(ctx, Ad, c["nb_alg"], c["pp"], c["panel_pivoting"], c["qr_type"], s)
    return rc, host(Ad), tau.cpu().numpy(), J.cpu().numpy(), s


@pytest.mark.parametrize("i", range(int(GH["hq_count"])))
def test_hqrrp_golden(ctx, i):
    A, st, c = hq_input(i)
    rc, F, tau, J, s = _run(ctx, A, c, st)
    check_hqrrp_against_golden(i, c, A, rc, F, tau, J, s.words())


@pytest.mark.parametrize("shape", [(1, 1, 1, 0), (7, 3, 2, 1), (40, 40, 8, 2), (100, 37, 16, 4), (37, 100, 16, 4), (300, 64, 64, 10), (300, 64, 100, 10),
                                   (50, 20, 64, 10)])
def test_hqrrp_edge_shapes_vs_oracle(ctx, shape):
    """Ragged blocks, a single block, wide inputs, a block size above both dimensions and a TALL uniform operator (nb_alg + pp > m):
    every panel-QR flavour against the restatement on the same input and state."""
    m, n, nb, pp = shape
    rng = np.random.default_rng(m * 31 + n)
    A = np.asfortranarray(rng.standard_normal((m, n)) * (1.0 + 0.3 * np.arange(n))[None, :] ** -1.0)
    for piv, qt in ((1, 0), (0, 0), (0, 1), (0, 2)):
        if qt == 2 and m < n:          # the CholQR panel needs m - j >= b in every block
            continue
        c = dict(nb_alg=nb, pp=pp, panel_pivoting=piv, qr_type=qt)
        rc, F, tau, J, s = _run(ctx, A, c, O.RNGState(0))
        rc2, F2, tau2, J2, st2 = O.hqrrp(A, nb, pp, piv, qt, O.RNGState(0))
        k = min(m, n)
        assert rc == rc2 and np.array_equal(J, J2), (shape, piv, qt)
        assert list(s.words()) == list(st2.words())
        assert np.abs(np.triu(F)[:k] - np.triu(F2)[:k]).max() <= 1e-9 * np.abs(F2).max(), (shape, piv, qt)
        assert np.abs(tau[:k] - tau2[:k]).max() <= 1e-9
        e = geqp3_format_invariants(A, F, tau, J, k)
        atol = np.finfo(np.float64).eps ** 0.75
        if qt == 2:
            # CholQR panels square the panel's condition number; the small trailing panels of these graded matrices are ill-conditioned
            # and the reference's own factors lose orthogonality to the same degree: measured against the restatement's figure
            e2 = geqp3_format_invariants(A, F2, tau2, J2, k)
            atol = max(atol, 10 * max(e2))
        assert max(e) <= atol, (shape, piv, qt, e)


def test_hqrrp_quick_return_host_call_and_timing(ctx):
    """min(m, n) = 0: nothing is written, the state does not move (rl_hqrrp.hh:886-888).  Host-pointer form = device form.  The nine leading
    entries of the reference's timing vector (:1140-1148) add up."""
    Z = rl.to_f(torch.zeros((0, 5), dtype=torch.float64, device="cuda"))
    s = rl.RNGState(3)
    J = torch.full((5,), -1, dtype=torch.int64, device="cuda")
    rc, tau, J = rl.hqrrp(ctx, Z, 4, 2, 1, 0, s, J=J)
    assert rc == 0 and list(s.words()) == list(rl.RNGState(3).words()) and J.cpu().tolist() == [-1] * 5
    A, st, c = hq_input(1)
    rc, F, tau, Jd, s1 = _run(ctx, A, c, st)
    Ah = torch.from_numpy(np.array(A.T, order="C", copy=True)).t()
    s2 = rl.RNGState(st.key, st.counter)
    ctx.check(ctx._lib.rlb200_set_phase_timing(ctx._h, 1))
    rc2, tau2, J2 = rl.hqrrp(ctx, Ah, c["nb_alg"], c["pp"], c["panel_pivoting"], c["qr_type"], s2)
    import ctypes
    buf = (ctypes.c_longlong * 32)()
    nt = ctx._lib.rlb200_get_phase_times(ctx._h, buf, 32)
    ctx.check(ctx._lib.rlb200_set_phase_timing(ctx._h, 0))
    assert rc2 == rc and np.array_equal(J2.numpy(), Jd) and list(s2.words()) == list(s1.words())
    assert np.abs(np.asfortranarray(Ah.numpy()) - F).max() <= 1e-12 and np.abs(tau2.numpy() - tau).max() <= 1e-12
    t = list(buf)[:nt]
    assert nt == 9 and all(x >= 0 for x in t) and sum(t[:8]) == t[8] and t[8] > 0


def test_hqrrp_engine_trailing_update_property(ctx):
    """A 20000 x 512 input: the compact-WY trailing update of the first blocks runs on the int8 digit-slice engine (rows - k >= 8192,
    apply_qt_wy).  The restatement on the same input and state: pivots exact, R to 1e-9; the reference's eps^0.75 acceptance test."""
    m, n, nb, pp = 20000, 512, 128, 16
    A, st = O.gen_poly_mat(m, n, n, 1e4, 2.0, O.RNGState(0))
    c = dict(nb_alg=nb, pp=pp, panel_pivoting=1, qr_type=0)
    rc, F, tau, J, s = _run(ctx, A, c, st)
    rc2, F2, tau2, J2, st2 = O.hqrrp(A, nb, pp, 1, 0, st)
    assert rc == rc2 == 0 and list(s.words()) == list(st2.words())
    assert np.array_equal(J, J2)
    assert np.abs(np.triu(F)[:n] - np.triu(F2)[:n]).max() <= 1e-9 * np.abs(np.diag(F2)).max()
    e = geqp3_format_invariants(A, F, tau, J, n)
    assert max(e) <= np.finfo(np.float64).eps ** 0.75, e


@pytest.mark.parametrize("shape", [(3000, 200, 1e3, 1.5, 64, 10, 1, 0), (2000, 120, 10.0, 1.25, 32, 8, 0, 1), (40000, 300, 1e2, 1.25, 64, 10, 1, 0)])
def test_cqrrpt_qrcp_hqrrp_vs_oracle(ctx, shape):
    """CQRRPT's `qrcp` field = hqrrp (rl_cqrrpt.hh:41, 230-231) with its HQRRP fields (:134-137): the device path against the restatement
    (pinned to the compiled reference in tests/test_oracle_qr.py::test_cqrrpt_qrcp_hqrrp) on the same input and state."""
    m, n, cond, df, nb, ov, piv, uc = shape
    A, st = O.gen_poly_mat(m, n, n, cond, 2.0, O.RNGState(0))
    alg = rl.CQRRPT(False, None)
    alg.qrcp = "hqrrp"
    alg.nb_alg, alg.oversampling, alg.panel_pivoting, alg.use_cholqr = nb, ov, piv, uc
    Ad = dev(A)
    s = rl.RNGState(st.key, st.counter)
    rc, R, J = alg.call(ctx, Ad, df, s)
    o = O.CQRRPT(float(np.finfo(np.float64).eps) ** 0.85, 2)
    o.qrcp = "hqrrp"
    o.nb_alg, o.oversampling, o.panel_pivoting, o.use_cholqr = nb, ov, piv, uc
    rc2, Q2, R2, J2, st2 = o.call(A, df, st)
    Jh, Rh, Qh = J.cpu().numpy(), host(R), host(Ad)
    assert (rc, alg.rank) == (rc2, o.rank) and list(s.words()) == list(st2.words())
    assert np.array_equal(Jh, J2)
    assert np.abs(np.triu(Rh) - np.triu(R2)).max() <= 1e-9 * np.abs(R2).max()
    e = qr_invariants(A, Qh, np.triu(Rh), Jh, alg.rank)
    assert max(e) <= np.finfo(np.float64).eps ** 0.75, e
    # the context setting is per object: the default (geqp3) is back for the next CQRRPT
    alg0 = rl.CQRRPT(False, None)
    rc0, _, J0 = alg0.call(ctx, dev(A), df, rl.RNGState(st.key, st.counter))
    assert rc0 == 0 and sorted(J0.cpu().tolist()) == list(range(1, n + 1))
