"""GPU parity, SYPS / SYRF / REVD2 through the C-ABI, against the goldens of the compiled reference and the oracle.

Tolerances (stated): return code, the final rank estimate k (the whole doubling sequence decides it) and the RNG state are exact;
eigenvalues 1e-10 of the largest one (1e-9 where the case says the truncated Ritz values are ill determined); the reconstruction error
||A - V E V^T||_F / ||A||_F within 1e-12 of the reference's; eigenvectors of well separated eigenvalues 1e-7 up to sign."""
import numpy as np
import pytest
import torch

import randlapack_b200 as rl
from _evdcases import CASES, GE, evd_matrix, check_revd2_against_golden
from oracle import rl_oracle as O

pytestmark = pytest.mark.gpu
UPLO = {0: "U", 1: "L"}
STAB = {1: rl.CholQRQ, 2: rl.HQRQ}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a.T)).cuda().t()


def host(t):
    return np.asfortranarray(t.cpu().numpy())


def _algs(c):
    syps = rl.SYPS(c["p"], c["q"])
    syrf = rl.SYRF(syps, STAB[c["orth"]]())
    return syps, syrf, rl.REVD2(syrf, c["est_p"])


def _state(st):
    return rl.RNGState((st[4], st[5]), tuple(st[:4]))


@pytest.mark.parametrize("i", range(len(CASES)))
def test_revd2_golden(ctx, i):
    c = CASES[i]
    A, S, st = evd_matrix(c)
    _, _, alg = _algs(c)
    s = _state(st)
    rc, k, V, ev = alg.call(ctx, UPLO[c["uplo"]], dev(A), c["k_start"], c["tol"], s)
    check_revd2_against_golden(i, c, S, rc, k, host(V), ev.cpu().numpy(), s.words(), eig_tol=max(1e-10, c.get("eig_tol", 0)))


@pytest.mark.parametrize("i", [0, 3, 5, 7])
def test_syps_syrf_vs_oracle(ctx, i):
    """Same operator, same state: the device sketch spans the oracle's subspace (leading columns agree up to sign), Q is orthonormal."""
    c = CASES[i]
    A, S, st = evd_matrix(c)
    ks = min(c["m"], 24)
    syps, syrf, _ = _algs(c)
    s = _state(st)
    rc, sk = syps.call(ctx, UPLO[c["uplo"]], dev(A), ks, s)
    assert rc == 0 and list(s.words()) == list(GE[f"ev{i}_syps_state"])
    h = GE[f"ev{i}_syps_head"]
    sk = host(sk)
    assert np.all(np.isfinite(sk))
    assert np.abs(np.abs(sk[:24, :6]) - np.abs(h[:, :6])).max() <= 1e-8 * max(1.0, np.abs(h).max())
    assert np.abs(np.abs(sk[:24]) - np.abs(h)).max() <= 1e-3 * max(1.0, np.abs(h).max())
    s = _state(st)
    rc, Q = syrf.call(ctx, UPLO[c["uplo"]], dev(A), ks, s)
    Q = host(Q)
    assert rc == 0 and list(s.words()) == list(GE[f"ev{i}_syrf_state"])
    assert np.abs(np.abs(Q[:24, :6]) - np.abs(GE[f"ev{i}_syrf_head"][:, :6])).max() <= 1e-8
    assert np.linalg.norm(Q.T @ Q - np.eye(ks)) <= 1e-12
    # captured energy equals the oracle's
    _, Qo, _ = O.SYRF(O.SYPS(c["p"], c["q"]), O.make_stab(c["orth"])).call(UPLO[c["uplo"]], A, ks, O.RNGState.from_words(st))
    e_dev, e_or = np.linalg.norm(Q.T @ S @ Q), np.linalg.norm(Qo.T @ S @ Qo)
    assert abs(e_dev - e_or) <= 1e-10 * e_or


def test_revd2_host_entry_and_capacity(ctx):
    """rlb200_revd2_f64_host == the device entry; k_cap below the converged rank -> code 3 with the last iterate."""
    c = CASES[1]
    A, S, st = evd_matrix(c)
    _, _, alg = _algs(c)
    s = _state(st)
    rc, k, V, ev = alg.call_host(ctx, "U", torch.from_numpy(np.ascontiguousarray(A.T)).t(), c["k_start"], c["tol"], s)
    check_revd2_against_golden(1, c, S, rc, k, np.asfortranarray(V.numpy()), ev.numpy(), s.words(), eig_tol=1e-10)
    s = _state(st)
    rc, k, V, ev = alg.call(ctx, "U", dev(A), c["k_start"], c["tol"], s, k_cap=20)
    assert rc == 3 and k == 20 and torch.isfinite(V).all()


def test_revd2_argument_checks(ctx):
    A, S, st = evd_matrix(CASES[3])
    _, _, alg = _algs(CASES[3])
    with pytest.raises(rl.Error):
        alg.call(ctx, "U", dev(A), 0, 1e-10, rl.RNGState(0))
    with pytest.raises(rl.Error):
        alg.call(ctx, "U", dev(A), 4, -1.0, rl.RNGState(0))


def test_revd2_fp32(ctx):
    """fp32 instantiation: PSD matrix of rank 30, residual at the fp32 level."""
    rng = np.random.default_rng(5)
    m, r = 512, 30
    U = np.linalg.qr(rng.standard_normal((m, r)))[0]
    lam = np.logspace(0, -3, r)
    S = ((U * lam) @ U.T).astype(np.float32)
    S = np.asfortranarray((S + S.T) / 2)
    alg = rl.REVD2(rl.SYRF(rl.SYPS(2, 1), rl.HQRQ()), 5)
    rc, k, V, ev = alg.call(ctx, "L", dev(S), 8, 1e-4, rl.RNGState(3))
    V, ev = host(V).astype(np.float64), ev.cpu().numpy().astype(np.float64)
    assert rc == 0 and k in (32, 64)
    assert np.linalg.norm(S - (V * ev) @ V.T) / np.linalg.norm(S) <= 5e-5
