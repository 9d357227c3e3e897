"""Shared inputs of the SYPS / SYRF / REVD2 parity tests: the matrices of test/drivers/test_revd2.cc, scaled down so the
oracle finishes in seconds, and the checks against the goldens of the compiled reference (tests/golden/make_golden_revd2.py)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_p = os.path.join(HERE, "golden", "revd2_vectors.npz")
GE = np.load(_p) if os.path.exists(_p) else None

STAB_HQRQ, STAB_CHOLQRQ = 2, 1
# test_revd2.cc: m = 1000, rank 100 / 159, cond 1e8 / 1e2, k_start 1 / 10, tol 1e-14, SYPS(3, 1), HQRQ, 10 error-estimate steps
CASES = [
    dict(m=1000, rank=100, cond=1e8, k_start=1, tol=1e-14, p=3, q=1, orth=STAB_HQRQ, est_p=10, uplo=0, nan=False),    # Underestimation1
    dict(m=1000, rank=100, cond=1e8, k_start=10, tol=1e-14, p=3, q=1, orth=STAB_HQRQ, est_p=10, uplo=0, nan=False),   # Underestimation2
    dict(m=1000, rank=100, cond=1e2, k_start=10, tol=1e-14, p=3, q=1, orth=STAB_HQRQ, est_p=10, uplo=0, nan=False),   # Overestimation1
    dict(m=100, rank=100, cond=1e2, k_start=10, tol=1e-14, p=3, q=1, orth=STAB_HQRQ, est_p=10, uplo=0, nan=False),    # Exactness (k reaches m)
    dict(m=100, rank=50, cond=1e2, k_start=1, tol=1e-14, p=3, q=1, orth=STAB_HQRQ, est_p=10, uplo=0, nan=True),       # Uplo (upper)
    dict(m=100, rank=50, cond=1e2, k_start=1, tol=1e-14, p=3, q=1, orth=STAB_HQRQ, est_p=10, uplo=1, nan=True),       # Uplo (lower)
    dict(m=600, rank=80, cond=1e4, k_start=16, tol=1e-10, p=2, q=2, orth=STAB_CHOLQRQ, est_p=5, uplo=1, nan=False),   # k outgrows the rank: the reference throws
    dict(m=400, rank=400, cond=1e4, k_start=16, tol=1e-5, p=2, q=2, orth=STAB_CHOLQRQ, est_p=5, uplo=1, nan=False, eig_tol=1e-9),   # even p, q = 2, CholQRQ; truncated at 1e-6: Ritz values move with rounding
    dict(m=400, rank=60, cond=1e3, k_start=8, tol=1e-6, p=0, q=1, orth=STAB_HQRQ, est_p=3, uplo=0, nan=False),        # no power passes
]


def evd_matrix(c, ref=None):
    """-> (A as handed to the algorithm [unused triangle NaN when c['nan']], the full symmetric matrix, input state words)."""
    m = c["m"]
    if ref is not None:
        import _ref
        B, st = _ref.ref_mat_gen(ref, 0, m, m, c["rank"], c["cond"], 2.0, [0] * 6, np.float64)
    else:
        from oracle import rl_oracle as O
        B, s = O.gen_poly_mat(m, m, c["rank"], c["cond"], 2.0, O.RNGState(0))
        st = list(s.words())
    S = B.T @ B                                   # syrk(Lower, Trans) + mirror (test_revd2.cc:103-115)
    S = np.asfortranarray(np.tril(S) + np.tril(S, -1).T)
    A = S.copy(order="F")
    if c["nan"]:
        iu = np.triu_indices(m, 1)
        if c["uplo"] == 0:
            A.T[iu] = np.nan                      # strictly lower triangle
        else:
            A[iu] = np.nan
    return A, S, st


def check_revd2_against_golden(i, c, S, rc, k, V, ev, state_words, eig_tol=1e-10, recon_slack=1e-12):
    g = GE
    if int(g[f"ev{i}_rc"][0]) != 0:          # the reference threw (std::runtime_error): the device path reports code 1 or 2 at the same k
        assert rc in (1, 2) and k == int(g[f"ev{i}_k"][0]), (rc, k)
        return
    assert rc == 0 and k == int(g[f"ev{i}_k"][0]), (rc, k, int(g[f"ev{i}_k"][0]))
    assert list(state_words) == list(g[f"ev{i}_state_out"])
    ev_ref = g[f"ev{i}_eig"]
    assert np.all(np.isfinite(ev)) and np.all(np.isfinite(V))
    assert np.abs(ev - ev_ref).max() <= eig_tol * ev_ref.max(), np.abs(ev - ev_ref).max() / ev_ref.max()
    recon = np.linalg.norm(S - (V * ev) @ V.T) / np.linalg.norm(S)
    assert recon <= float(g[f"ev{i}_recon"][0]) + recon_slack, (recon, float(g[f"ev{i}_recon"][0]))
    # eigenvectors of well separated eigenvalues agree up to sign (first rows of V are in the golden file)
    Vh = g[f"ev{i}_Vhead"]
    gaps = np.minimum(np.abs(np.diff(ev_ref, prepend=np.inf)), np.abs(np.diff(ev_ref, append=-np.inf))) / ev_ref.max()
    sel = (gaps > 1e-3) & (ev_ref > 1e-6 * ev_ref.max())
    if sel.any():
        d = np.abs(np.abs(V[:24, sel]) - np.abs(Vh[:, sel])).max()
        assert d <= 1e-7, d
