"""GPU parity of the stabiliser / rangefinder / QB / RSVD drivers against the oracle on the same inputs.

Parity rule (north star): subspace angle and ||A - QQ'A||_F residual within 1e-10 of the reference's (fp64),
singular values to 1e-10 relative; return codes and RNG state identical.  The reference's own invariants for
this path are re-asserted with the reference's exponents (test_orth.cc:98, test_rf.cc:132,139, test_qb.cc:162-174).

The Gaussian operator entries differ from the host libm path in at most a couple of float ulps (test_gpu_fill.py),
which moves a randomized result by ~1e-7 relative.  To compare at 1e-10 the oracle is therefore run on the very
operator the device generated (`omega_override`), exactly as the reference's own CPU-vs-GPU test feeds one
host-built sketch to both paths (test/drivers/test_bqrrp_gpu.cu:91-103); a second comparison with each side's own
operator is made at the looser tolerance that the ulp difference justifies."""
import os

import numpy as np
import pytest
import torch

import _ref
import randlapack_b200 as rl
from oracle import rl_oracle as O

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a.T)).cuda().t()
    return t if dtype is None else t.to(dtype)


def host(t):
    return np.asfortranarray(t.detach().cpu().numpy())


def poly(m, n, rank, cond=2025.0, expo=2.0, dtype=np.float64):
    return O.gen_poly_mat(m, n, rank, cond, expo, O.RNGState(0), dtype=dtype)


# ---------------------------------------------------------------- stabilisers
def test_cholqrq_matches_oracle_and_reference_invariant(ctx):
    # test/comps/test_orth.cc:135-153
    m, n, k = 1000, 200, 200
    A, st = poly(m, n, k, 2.0)
    Om, _ = O.fill_dense(n, k, st)
    Y = np.asfortranarray(A @ Om)
    rc_o, Q_o = O.CholQRQ().call(Y.copy(order="F"))
    Yd = dev(Y).clone()
    orth = rl.CholQRQ(False, False)
    rc = orth.call(ctx, Yd)
    assert rc == rc_o == 0 and not orth.chol_fail
    Q = host(Yd)
    # same factorisation (unique: R has a positive diagonal) => entries agree to cond(Y)^2 * eps
    assert np.abs(Q - Q_o).max() <= 1e-9
    rc = orth.call(ctx, Yd)
    Q = host(Yd)
    assert rc == 0 and np.linalg.norm(Q.T @ Q - np.eye(k)) <= EPS ** 0.625


@pytest.mark.parametrize("m,k", [(50, 4), (5000, 37), (100000, 256), (70, 70)])
def test_cholqrq_shapes(ctx, m, k):
    rng = np.random.default_rng(m)
    Y = np.asfortranarray(rng.standard_normal((m, k)))
    rc_o, Q_o = O.CholQRQ().call(Y.copy(order="F"))
    Yd = dev(Y).clone()
    rc = rl.CholQRQ().call(ctx, Yd)
    assert rc == rc_o == 0
    assert np.abs(host(Yd) - Q_o).max() <= 1e-10 * max(1.0, np.linalg.cond(Y) ** 2)


def test_cholqrq_failure_code(ctx):
    # rl_orth.hh:81-85 — rank-deficient input: potrf fails => return 1, chol_fail set, same as the oracle
    Y = np.asfortranarray(np.ones((50, 4)))
    o = O.CholQRQ()
    rc_o, _ = o.call(Y.copy(order="F"))
    orth = rl.CholQRQ()
    rc = orth.call(ctx, dev(Y).clone())
    assert rc == rc_o == 1 and orth.chol_fail and o.chol_fail


def test_cholqrq_cond_check(ctx):
    # rl_orth.hh:88-93 — cond(R) > 1/sqrt(eps) => return 1
    rng = np.random.default_rng(3)
    U, _ = np.linalg.qr(rng.standard_normal((400, 8)))
    Y = np.asfortranarray(U * np.logspace(0, -9, 8))
    rc_o, _ = O.CholQRQ(cond_check=True).call(Y.copy(order="F"))
    rc = rl.CholQRQ(cond_check=True).call(ctx, dev(Y).clone())
    assert rc == rc_o == 1
    Y = np.asfortranarray(U * np.logspace(0, -3, 8))
    assert rl.CholQRQ(cond_check=True).call(ctx, dev(Y).clone()) == 0


def test_svd_tall(ctx):
    rng = np.random.default_rng(0)
    for (n, k) in [(1024, 256), (300, 7), (64, 64), (50, 1), (257, 33)]:
        B = np.asfortranarray(rng.standard_normal((n, k)) * np.logspace(0, -6, k))
        Bd = dev(B).clone()
        V, S, W = rl.svd_tall(ctx, Bd)
        V, S, W = host(V), S.cpu().numpy(), host(W)
        s_ref = np.linalg.svd(B, compute_uv=False)
        assert np.all(np.diff(S) <= 0)
        assert np.abs(S - s_ref).max() <= 1e-13 * s_ref[0]
        assert np.linalg.norm(V.T @ V - np.eye(k)) <= 1e-12 * k
        assert np.linalg.norm(W.T @ W - np.eye(k)) <= 1e-12 * k
        assert np.linalg.norm((V * S) @ W.T - B) <= 1e-13 * k * s_ref[0]


# ---------------------------------------------------------------- RS / RF / QB / RSVD
def _stack(p, q, b, stab=rl.STAB_CHOLQRQ, orth_check=False):
    cls = {rl.STAB_CHOLQRQ: rl.CholQRQ, rl.STAB_PLUL: rl.PLUL, rl.STAB_HQRQ: rl.HQRQ}[stab]
    Stab = cls(False, False)
    RS = rl.RS(Stab, p, q, False, False)
    RF = rl.RF(RS, rl.CholQRQ(False, False), False, False)
    QB = rl.QB(RF, rl.CholQRQ(False, False), False, orth_check)
    return RS, RF, QB, rl.RSVD(QB, b)


def _device_operator(ctx, m, n, k, p, st, dtype=torch.float64):
    """The operator the device will draw for RS (rl_rs.hh:132-139), as a host array."""
    rows = n if p % 2 == 0 else m
    buf, _ = rl.fill_dense(ctx, rl.DenseDist(rows, k), st.copy(), dtype)
    # the reference hands the natural-layout buffer to BLAS as a column-major rows x k matrix (rl_rs.hh:135,139,153)
    return np.asfortranarray(buf.cpu().numpy().astype(np.float64).reshape((rows, k), order="F"))


@pytest.mark.parametrize("m,n,k,p,q", [(400, 64, 16, 0, 1), (400, 64, 16, 2, 1), (400, 64, 16, 3, 1), (2000, 300, 40, 4, 2), (1000, 100, 8, 1, 1)])
def test_rs_rf_vs_oracle(ctx, m, n, k, p, q):
    A, st0 = poly(m, n, n)
    Ad = dev(A)
    st_d = rl.RNGState(st0.key, st0.counter)
    Om_dev = _device_operator(ctx, m, n, k, p, st_d)
    RS, RF, _, _ = _stack(p, q, k)
    o = O.StackOpts(p, q, k, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ)
    rs_o, rf_o, _, _ = O.make_stack(o)
    # RS
    s1 = st_d.copy()
    rc, Om = RS.call(ctx, Ad, k, s1)
    rc_o, Om_o, s1_o = rs_o.call(A, k, st0.copy(), omega_override=Om_dev)
    assert rc == rc_o == 0 and s1.counter == s1_o.counter and s1.key == s1_o.key
    scale = np.abs(Om_o).max()
    assert np.abs(host(Om) - Om_o).max() <= 1e-9 * scale
    # RF: orthonormal basis of the same subspace
    s2 = st_d.copy()
    rc, Q = RF.call(ctx, Ad, k, s2)
    rc_o, Q_o, s2_o = rf_o.call(A, k, st0.copy(), omega_override=Om_dev)
    Q = host(Q)
    assert rc == rc_o == 0 and s2.counter == s2_o.counter
    assert np.linalg.norm(Q.T @ Q - np.eye(k)) <= EPS ** 0.625                     # test_rf.cc:132
    assert _ref.subspace_sin(Q_o, Q) <= 1e-9
    assert np.abs(Q - Q_o).max() <= 1e-8


@pytest.mark.parametrize("i", range(int(GOLD["rsvd_count"])))
def test_rsvd_golden_configs(ctx, i):
    """The reference's outputs (golden, generated from the compiled reference) on its own inputs, incl. BASELINE configs[0]."""
    m, n, k, p, q, b, stab = [int(x) for x in GOLD[f"rsvd{i}_args"]]
    cond, expo = GOLD[f"rsvd{i}_cond_expo"]
    A, st0 = poly(m, n, n if m > 10 else k, cond, expo)
    *_, RSVD = _stack(p, q, b, stab=stab)
    st = rl.RNGState(st0.key, st0.counter)
    rc, kk, U, S, V = RSVD.call(ctx, dev(A), k, 0.0, st)
    assert [rc, kk] == [int(x) for x in GOLD[f"rsvd{i}_rc_k"]]
    assert list(st.counter) + list(st.key) == [int(x) for x in GOLD[f"rsvd{i}_state_out"]]
    U, S, V = host(U), S.cpu().numpy(), host(V)
    # own-operator comparison: Gaussian entries differ by <= 2 float ulps => results agree to ~1e-6 relative
    assert np.allclose(S, GOLD[f"rsvd{i}_S"], rtol=2e-6, atol=1e-9)
    resid = np.linalg.norm(A - (U * S) @ V.T) / np.linalg.norm(A)
    assert abs(resid - GOLD[f"rsvd{i}_resid"][0]) <= 2e-6
    assert np.linalg.norm(U.T @ U - np.eye(kk)) <= EPS ** 0.625 and np.linalg.norm(V.T @ V - np.eye(kk)) <= EPS ** 0.625


@pytest.mark.parametrize("m,n,k,p,q,b", [(4096, 256, 32, 0, 1, 32), (4096, 256, 32, 2, 1, 32), (1500, 200, 24, 2, 1, 8), (1500, 200, 24, 3, 1, 6),
                                          (600, 600, 20, 1, 1, 20), (3000, 64, 48, 0, 1, 48)])
def test_qb_rsvd_vs_oracle_same_operator(ctx, m, n, k, p, q, b):
    A, st0 = poly(m, n, n)
    Ad = dev(A)
    st_d = rl.RNGState(st0.key, st0.counter)
    _, _, QB, RSVD = _stack(p, q, b, orth_check=True)
    o = O.StackOpts(p, q, b, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ, False, True)
    _, _, qb_o, rsvd_o = O.make_stack(o)
    single = b == k
    Om_dev = _device_operator(ctx, m, n, k, p, st_d) if single else None
    # --- QB
    s = st_d.copy()
    rc, kk, Q, BT = QB.call(ctx, Ad, k, b, 0.0, s)
    rc_o, kk_o, Q_o, BT_o, s_o = qb_o.call(A, k, b, 0.0, st0.copy(), omega_override=Om_dev)
    assert (rc, kk) == (rc_o, kk_o) and s.counter == s_o.counter and s.key == s_o.key
    if kk == 0:      # e.g. orthogonality_check tripped (rc 4): identical failure on both sides, nothing more to compare
        return
    Q, BT = host(Q)[:, :kk], host(BT)[:, :kk]
    nrmA = np.linalg.norm(A)
    r_dev = np.linalg.norm(A - Q @ (Q.T @ A)) / nrmA
    r_ora = np.linalg.norm(A - Q_o @ (Q_o.T @ A)) / nrmA
    tol = 1e-10 if single else 2e-6     # multi-block: every block draws its own operator (ulp-level differences)
    assert abs(r_dev - r_ora) <= tol
    assert _ref.subspace_sin(Q_o, Q) <= (1e-9 if single else 1e-4)
    assert np.linalg.norm(Q.T @ Q - np.eye(kk)) <= EPS ** 0.625                      # test_qb.cc:166
    assert np.linalg.norm(A - Q @ BT.T) <= np.linalg.norm(A - Q_o @ BT_o.T) + 1e-9 * nrmA
    # --- RSVD
    s = st_d.copy()
    rc, kk, U, S, V = RSVD.call(ctx, Ad, k, 0.0, s)
    rc_o, kk_o, U_o, S_o, V_o, s_o = rsvd_o.call(A, k, 0.0, st0.copy(), omega_override=Om_dev)
    assert (rc, kk) == (rc_o, kk_o) and s.counter == s_o.counter
    U, S, V = host(U)[:, :kk], S.cpu().numpy()[:kk], host(V)[:, :kk]
    assert np.abs(S - S_o).max() <= (1e-10 if single else 2e-6) * S_o[0]
    assert abs(np.linalg.norm(A - (U * S) @ V.T) - np.linalg.norm(A - (U_o * S_o) @ V_o.T)) <= tol * nrmA
    assert _ref.subspace_sin(U_o, U) <= (1e-9 if single else 1e-4)
    assert np.linalg.norm(U.T @ U - np.eye(kk)) <= EPS ** 0.625 and np.linalg.norm(V.T @ V - np.eye(kk)) <= EPS ** 0.625


@pytest.mark.parametrize("digits", [6, 7, -1])
@pytest.mark.parametrize("m,n,k,p", [(32768, 256, 32, 0), (32768, 256, 32, 2), (20000, 300, 40, 3)])
def test_rsvd_i8_engine_vs_oracle(ctx, m, n, k, p, digits):
    """The same parity rule on each engine of the tall products over A — the tcgen05 int8 digit-slice engine (ozaki.cu, the
    default; 6 and 7 digits) and the fp64 DMMA pipe (digits = -1): subspace angle, residual and singular values within
    1e-10 / 1e-9 of the oracle's on the same operator, identical codes and RNG state."""
    A, st0 = poly(m, n, n)
    Ad = dev(A)
    st_d = rl.RNGState(st0.key, st0.counter)
    _, _, QB, RSVD = _stack(p, 1, k)
    o = O.StackOpts(p, 1, k, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ)
    *_, rsvd_o = O.make_stack(o)
    Om_dev = _device_operator(ctx, m, n, k, p, st_d)
    ctx.set_fp64_engine("i8" if digits > 0 else "dmma")
    ctx.set_i8_digits(max(digits, 0))
    try:
        s = st_d.copy()
        rc, kk, U, S, V = RSVD.call(ctx, Ad, k, 0.0, s)
    finally:
        ctx.set_fp64_engine("i8")
        ctx.set_i8_digits(0)
    rc_o, kk_o, U_o, S_o, V_o, s_o = rsvd_o.call(A, k, 0.0, st0.copy(), omega_override=Om_dev)
    assert (rc, kk) == (rc_o, kk_o) and s.counter == s_o.counter and s.key == s_o.key
    U, S, V = host(U), S.cpu().numpy(), host(V)
    nrmA = np.linalg.norm(A)
    assert np.abs(S - S_o).max() <= 1e-10 * S_o[0]
    assert abs(np.linalg.norm(A - (U * S) @ V.T) - np.linalg.norm(A - (U_o * S_o) @ V_o.T)) <= 1e-10 * nrmA
    assert _ref.subspace_sin(U_o, U) <= 1e-9
    assert np.linalg.norm(U.T @ U - np.eye(kk)) <= EPS ** 0.625 and np.linalg.norm(V.T @ V - np.eye(kk)) <= EPS ** 0.625


@pytest.mark.parametrize("graded", [False, True])
@pytest.mark.parametrize("k,p", [(64, 0), (64, 3), (128, 2), (256, 0), (256, 2), (256, 3)])
def test_rsvd_headline_engine_path_vs_oracle(ctx, k, p, graded):
    """The code path every BASELINE config takes on the int8 engine (k >= 64, m >= 16384): Gram matrix fused into the A^T Y launch,
    R folded into the next product inside RS, U = Y (R^-1 W) on the engine (drivers.cu: tall_tn_gram / rs_call / rsvd_single_block_fused).
    Planted sigma_i = i^-2 (gen_poly_mat, cond 2025) and a column-graded variant (columns scaled over 12 decades: the worst case of the
    per-row digit scaling).  Same-operator rule: sigma 1e-10 relative to sigma_1, residual within 1e-10, subspace sin 1e-9, codes/state equal."""
    m, n = 32768, 512
    A, st0 = poly(m, n, n)
    if graded:
        A = np.asfortranarray(A * np.logspace(0, -12, n)[None, :])
    Ad = dev(A)
    st_d = rl.RNGState(st0.key, st0.counter)
    *_, RSVD = _stack(p, 1, k)
    o = O.StackOpts(p, 1, k, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ)
    *_, rsvd_o = O.make_stack(o)
    Om_dev = _device_operator(ctx, m, n, k, p, st_d)
    s = st_d.copy()
    rc, kk, U, S, V = RSVD.call(ctx, Ad, k, 0.0, s)
    rc_o, kk_o, U_o, S_o, V_o, s_o = rsvd_o.call(A, k, 0.0, st0.copy(), omega_override=Om_dev)
    assert (rc, kk) == (rc_o, kk_o) and s.counter == s_o.counter and s.key == s_o.key
    if kk == 0:
        return
    U, S, V = host(U)[:, :kk], S.cpu().numpy()[:kk], host(V)[:, :kk]
    nrmA = np.linalg.norm(A)
    assert np.abs(S - S_o).max() <= 1e-10 * S_o[0]
    assert abs(np.linalg.norm(A - (U * S) @ V.T) - np.linalg.norm(A - (U_o * S_o) @ V_o.T)) <= 1e-10 * nrmA
    # subspaces: compare the part of the basis whose singular values are resolved at 1e-10 (directions below that are round-off
    # in the reference as well); with the graded matrix the trailing directions carry sigma ~ 1e-12 sigma_1
    r = int(np.sum(S_o > 1e-6 * S_o[0])) if graded else kk
    assert _ref.subspace_sin(U_o[:, :r], U[:, :r]) <= (1e-6 if graded else 1e-9)
    # orthogonality: the reference's own tolerance, or - where CholQR of the graded sketch loses orthogonality in the reference itself
    # (cond(Y)^2 eps) - no worse than a small multiple of what the reference delivers on the same input
    o_ref = max(np.linalg.norm(U_o.T @ U_o - np.eye(kk)), np.linalg.norm(V_o.T @ V_o - np.eye(kk)))
    lim = max(EPS ** 0.625, 10 * o_ref)
    assert np.linalg.norm(U.T @ U - np.eye(kk)) <= lim and np.linalg.norm(V.T @ V - np.eye(kk)) <= lim


def test_rsvd_nonfinite_input_propagates(ctx):
    """a NaN / Inf entry must not come back as plausible finite factors.  The reference's BLAS path propagates the NaN into the
    Gram matrix, potrf fails, RF returns 2 and QB/RSVD report k = 0 (rl_qb.hh:191-197).  Both engines must do the same."""
    m, n, k = 20000, 128, 64
    A, st0 = poly(m, n, n)
    for bad in (np.nan, np.inf):
        Ab = A.copy(order="F")
        Ab[1234, 17] = bad
        for engine in ("dmma", "i8"):
            ctx.set_fp64_engine(engine)
            try:
                *_, RSVD = _stack(2, 1, k)
                rc, kk, U, S, V = RSVD.call(ctx, dev(Ab), k, 0.0, rl.RNGState(st0.key, st0.counter))
            finally:
                ctx.set_fp64_engine("i8")
            assert kk == 0 or not np.isfinite(S.cpu().numpy()).all(), (bad, engine, rc, kk)


def test_rsvd_f32(ctx):
    m, n, k = 2000, 128, 16
    A, st0 = poly(m, n, n, dtype=np.float32)
    *_, RSVD = _stack(2, 1, k)
    o = O.StackOpts(2, 1, k, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ)
    *_, rsvd_o = O.make_stack(o)
    st = rl.RNGState(st0.key, st0.counter)
    rc, kk, U, S, V = RSVD.call(ctx, dev(A), k, 0.0, st)
    rc_o, kk_o, U_o, S_o, V_o, s_o = rsvd_o.call(A, k, 0.0, st0.copy())
    assert (rc, kk) == (rc_o, kk_o) and st.counter == s_o.counter
    assert np.abs(S.cpu().numpy() - S_o).max() <= 1e-4 * S_o[0]


def test_rsvd_f32_tall_on_the_digit_slice_engine(ctx):
    """fp32 storage, tall enough (m >= 16384) for the tall products to run on the int8 digit-slice engine with 4 digits (30 bits):
    singular values against the fp64 oracle on the same fp32 input, orthonormal factors."""
    m, n, k = 40000, 128, 16
    A, st0 = poly(m, n, n, dtype=np.float32)
    *_, RSVD = _stack(2, 1, k)
    o = O.StackOpts(2, 1, k, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ)
    *_, rsvd_o = O.make_stack(o)
    st = rl.RNGState(st0.key, st0.counter)
    rc, kk, U, S, V = RSVD.call(ctx, dev(A), k, 0.0, st)
    rc_o, kk_o, U_o, S_o, V_o, s_o = rsvd_o.call(A, k, 0.0, st0.copy())
    assert (rc, kk) == (rc_o, kk_o) and st.counter == s_o.counter
    assert np.abs(S.cpu().numpy() - S_o).max() <= 1e-4 * S_o[0]
    Uh = host(U).astype(np.float64)
    assert np.linalg.norm(Uh.T @ Uh - np.eye(kk)) <= 1e-3


def test_rsvd_host_entry_point(ctx):
    # the reference-facing form: host buffers in, host buffers out (what e2e measures)
    m, n, k = 3000, 128, 16
    A, st0 = poly(m, n, n)
    *_, RSVD = _stack(0, 1, k)
    st = rl.RNGState(st0.key, st0.counter)
    Ah = torch.from_numpy(np.ascontiguousarray(A.T)).t()
    rc, kk, U, S, V = RSVD.call_host(ctx, Ah, k, 0.0, st)
    st2 = rl.RNGState(st0.key, st0.counter)
    rc2, kk2, U2, S2, V2 = RSVD.call(ctx, dev(A), k, 0.0, st2)
    assert (rc, kk) == (rc2, kk2) and st == st2
    assert torch.equal(S, S2.cpu()) and torch.equal(U, U2.cpu()) and torch.equal(V, V2.cpu())


def test_rsvd_argument_errors(ctx):
    # rl_rsvd.hh:128-132 — bad arguments are reported (RandLAPACK::Error in the reference), never UB
    A = dev(np.asfortranarray(np.ones((10, 5))))
    *_, RSVD = _stack(0, 1, 2)
    with pytest.raises(rl.Error):
        RSVD.call(ctx, A, 0, 0.0, rl.RNGState())
    with pytest.raises(rl.Error):
        RSVD.call(ctx, A, 2, -1.0, rl.RNGState())


@pytest.mark.parametrize("m,k", [(50, 5), (1000, 37), (300, 300), (20000, 64), (7, 7), (1, 1)])
def test_plul_vs_oracle(ctx, m, k):
    """PLUL::call (rl_orth.hh:211-230): same pivots => L agrees with the oracle's getrf/get_L/laswp to round-off."""
    rng = np.random.default_rng(m * 7 + k)
    Y = np.asfortranarray(rng.standard_normal((m, k)))
    rc_o, L_o = O.PLUL().call(Y.copy(order="F"))
    Yd = dev(Y).clone()
    rc = rl.PLUL().call(ctx, Yd)
    assert rc == rc_o == 0
    L = host(Yd)
    assert np.abs(L).max() <= 1.0 + 1e-12          # partial pivoting: |l_ij| <= 1
    assert np.abs(L - L_o).max() <= 1e-10


def test_plul_zero_column_and_ties(ctx):
    # a zero pivot column is tolerated (rl_orth.hh:218-222); ties resolve to the first row, as idamax does
    Y = np.asfortranarray(np.array([[2.0, 0.0, 1.0], [-2.0, 0.0, 3.0], [1.0, 0.0, -3.0], [2.0, 0.0, 0.5]]))
    rc_o, L_o = O.PLUL().call(Y.copy(order="F"))
    Yd = dev(Y).clone()
    assert rl.PLUL().call(ctx, Yd) == rc_o == 0
    assert np.allclose(host(Yd), L_o, atol=1e-14)


def test_plul_f32(ctx):
    rng = np.random.default_rng(5)
    Y = np.asfortranarray(rng.standard_normal((500, 20)).astype(np.float32))
    rc_o, L_o = O.PLUL().call(Y.copy(order="F"))
    Yd = dev(Y).clone()
    assert rl.PLUL().call(ctx, Yd) == 0
    assert np.abs(host(Yd) - L_o).max() <= 1e-4


def test_canonical_stack_with_plul(ctx):
    """The reference's canonical RSVD stack (test/drivers/test_rsvd.cc:68-93): PLUL stabiliser, CholQRQ orthogonalisers."""
    m, n, k, p = 1500, 200, 24, 2
    A, st0 = poly(m, n, n)
    _, _, QB, RSVD = _stack(p, 1, k, stab=rl.STAB_PLUL)
    o = O.StackOpts(p, 1, k, O.STAB_PLUL, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ)
    *_, rsvd_o = O.make_stack(o)
    st_d = rl.RNGState(st0.key, st0.counter)
    Om_dev = _device_operator(ctx, m, n, k, p, st_d)
    s = st_d.copy()
    rc, kk, U, S, V = RSVD.call(ctx, dev(A), k, 0.0, s)
    rc_o, kk_o, U_o, S_o, V_o, s_o = rsvd_o.call(A, k, 0.0, st0.copy(), omega_override=Om_dev)
    assert (rc, kk) == (rc_o, kk_o) and s.counter == s_o.counter
    assert np.abs(S.cpu().numpy() - S_o).max() <= 1e-10 * S_o[0]
    assert _ref.subspace_sin(U_o, host(U)) <= 1e-9


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(50, 5), (2000, 64), (257, 256), (1, 1), (4096, 33)])
def test_hqrq_vs_oracle(ctx, dtype, shape):
    """HQRQ::call (rl_orth.hh:144-164, geqrf + ungqr): Householder QR is unique, so Q equals the oracle's to round-off —
    also on an ill-conditioned iterate (cond 1e10), which is what HQRQ exists for."""
    m, k = shape
    rng = np.random.default_rng(m + k)
    A = np.asfortranarray((rng.standard_normal((m, k)) * np.logspace(0, -10 if dtype == np.float64 else -4, k)).astype(dtype))
    Ad = dev(A).clone()
    rc = rl.HQRQ(False, False).call(ctx, Ad)
    rc_o, Qo = O.HQRQ(False).call(A.copy(order="F"))
    Q = host(Ad)
    tol = 1e-12 if dtype == np.float64 else 1e-4
    assert rc == rc_o == 0
    assert np.linalg.norm(Q.T.astype(np.float64) @ Q - np.eye(k)) <= tol * k
    assert np.abs(Q - Qo).max() <= tol * 10


def test_folded_solves_match_explicit(ctx):
    """CholQR's triangular solve folded into the next product (default) vs the reference's explicit order of operations:
    identical return codes / k / RNG state, U, S, V equal to round-off."""
    A, st0 = poly(3000, 200, 200)
    for p in (0, 2, 3):
        outs = []
        for fold in (True, False):
            RS, RF, QB, RSVD = _stack(p, 1, 24)
            RS.fold_solves = fold
            s = rl.RNGState(st0.key, st0.counter)
            rc, kk, U, S, V = RSVD.call(ctx, dev(A), 24, 0.0, s)
            outs.append((rc, kk, RSVD.qb_code, s.counter, host(U), S.cpu().numpy(), host(V)))
        a, b = outs
        assert a[:4] == b[:4]
        assert np.abs(a[5] - b[5]).max() <= 1e-12 * a[5][0]
        assert _ref.subspace_sin(a[4], b[4]) <= 1e-9 and _ref.subspace_sin(a[6], b[6]) <= 1e-9
        Ur = (a[4] * a[5]) @ a[6].T - (b[4] * b[5]) @ b[6].T
        assert np.linalg.norm(Ur) <= 1e-11 * a[5][0]


def test_rsvd_large_properties(ctx):
    """Size-independent properties at a size the oracle cannot reach in seconds (2^22 x 512, k = 64):
    planted low-rank + noise => recovered spectrum, orthonormal factors, residual at the noise floor."""
    m, n, r, k = 1 << 22, 512, 32, 32
    g = torch.Generator(device="cuda").manual_seed(1)
    st = rl.RNGState(77)
    G1, st = rl.fill_dense(ctx, rl.DenseDist(m, r), st)
    G1 = G1.view(r, m).t()
    G2 = torch.randn((n, r), dtype=torch.float64, device="cuda", generator=g)
    sig = torch.logspace(0, -2, r, dtype=torch.float64, device="cuda")
    A = rl.empty_f(m, n, torch.float64, "cuda")
    A.copy_((G1 * sig) @ G2.t())
    *_, RSVD = _stack(2, 1, k)
    rc, kk, U, S, V = RSVD.call(ctx, A, k, 0.0, rl.RNGState(0))
    assert rc == 0 and kk == k
    I = torch.eye(k, dtype=torch.float64, device="cuda")
    assert torch.linalg.norm(U.t() @ U - I).item() <= 1e-9 and torch.linalg.norm(V.t() @ V - I).item() <= 1e-9
    R = A - (U * S) @ V.t()
    assert torch.linalg.norm(R).item() <= 1e-9 * torch.linalg.norm(A).item()
    AtA = A.t() @ A
    ev = torch.linalg.eigvalsh(AtA).flip(0)[:r].clamp_min(0).sqrt()
    assert torch.allclose(S[:8], ev[:8], rtol=1e-8)


def test_rank_deficient_failure_codes_match(ctx):
    """CholQRQ on a numerically rank-deficient sketch: potrf fails, RF returns 2, QB returns 6 with k = 0
    (rl_rf.hh:129-132, rl_qb.hh:191-197) — identical codes from the device path and the oracle."""
    m, n, k = 500, 40, 40
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.standard_normal((m, 8)) @ rng.standard_normal((8, n)))   # rank 8 < k
    _, RF, QB, _ = _stack(0, 1, k)
    o = O.StackOpts(0, 1, k, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ)
    _, rf_o, qb_o, _ = O.make_stack(o)
    rc, _ = RF.call(ctx, dev(A), k, rl.RNGState(0))
    rc_o, _, _ = rf_o.call(A, k, O.RNGState(0))
    assert rc == rc_o == 2
    rc, kk, _, _ = QB.call(ctx, dev(A), k, k, 0.0, rl.RNGState(0))
    rc_o, kk_o, *_ = qb_o.call(A, k, k, 0.0, O.RNGState(0))
    assert (rc, kk) == (rc_o, kk_o) == (6, 0)
