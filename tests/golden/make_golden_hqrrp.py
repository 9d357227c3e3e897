"""Generate tests/golden/hqrrp_vectors.npz from the REAL reference (oracle/_ref/librl_ref.so): RandLAPACK::hqrrp (rl_hqrrp.hh:811-1196) on
mat_gen inputs of the kind its own test uses (test/drivers/test_hqrrp.cc:153-172: 500 x 200 polynomial decay, b_sz 50, d_factor 1,
use_cholqr 1, no panel pivoting) plus the CQRRPT configuration (nb_alg 64, oversampling 10, panel pivoting, rl_cqrrpt.hh:60-63),
wide / square / ragged shapes and fp32.  Run in the build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _ref  # noqa: E402

R = _ref.ref_lib()
assert R is not None
R.rlref_set_num_threads(1)
out = {}
# (m, n, rank, cond, nb_alg, pp, panel_pivoting, qr_type, dtype)
CASES = [
    (500, 200, 200, 2.0, 50, 50, 0, 1, "f64"),       # test_hqrrp.cc HQRRP_full_rank_cholqr (use_cholqr = 1 selects geqrf, :582)
    (500, 200, 200, 100.0, 50, 50, 1, 0, "f64"),     # panel pivoting
    (2000, 600, 600, 1e4, 64, 10, 1, 0, "f64"),      # CQRRPT's HQRRP configuration
    (2000, 600, 600, 1e4, 64, 10, 0, 2, "f64"),      # CholQR + Householder reconstruction panels
    (300, 300, 300, 100.0, 64, 10, 1, 0, "f64"),     # square
    (200, 333, 200, 100.0, 32, 5, 1, 0, "f64"),      # wide: the last block ends at row m
    (257, 129, 129, 50.0, 64, 10, 0, 0, "f64"),      # ragged last block, unblocked Householder panel without pivoting
    (1500, 130, 130, 100.0, 64, 10, 1, 0, "f32"),
    (1000, 300, 120, 1e3, 64, 10, 1, 0, "f64"),      # rank-deficient: pivots past the rank are round-off
]
for i, (m, n, rk, cond, nb, pp, piv, qt, dt) in enumerate(CASES):
    npdt = np.float64 if dt == "f64" else np.float32
    A, st = _ref.ref_mat_gen(R, 0, m, n, rk, cond, 2.0, [0] * 6, npdt)
    rc, F, tau, J, st2 = _ref.ref_hqrrp(R, A, nb, pp, piv, qt, st)
    k = min(m, n)
    out[f"hq{i}_args"] = np.array([m, n, rk, nb, pp, piv, qt], dtype=np.int64)
    out[f"hq{i}_fargs"] = np.array([cond, 2.0])
    out[f"hq{i}_dtype"] = np.array(dt)
    out[f"hq{i}_state_in"], out[f"hq{i}_state_out"] = np.array(st, dtype=np.uint32), np.array(st2, dtype=np.uint32)
    out[f"hq{i}_rc"] = np.array([rc], dtype=np.int64)
    out[f"hq{i}_J"] = J
    out[f"hq{i}_tau"] = tau
    out[f"hq{i}_Rdiag"] = np.diag(F)[:k].copy()
    out[f"hq{i}_Fhead"] = F[:48, :48].copy()
    out[f"hq{i}_Achk"] = np.array([A.sum(dtype=np.float64), np.abs(A).sum(dtype=np.float64), A[0, 0], A[-1, -1]])
    print("hqrrp", i, (m, n), "rc", rc, "|Rdiag| head", np.abs(np.diag(F))[:3], "tail", np.abs(np.diag(F))[k - 1])
out["hq_count"] = np.array(len(CASES))
np.savez_compressed(os.path.join(HERE, "hqrrp_vectors.npz"), **out)
