"""Generate tests/golden/cqrrt_vectors.npz from the REAL reference (oracle/_ref/librl_ref.so): CQRRT (rl_cqrrt.hh:91-297) on mat_gen
inputs of the kind its own tests use (test/drivers/test_cqrrt.cc: polynomial decay, full rank).  Run in the build container only."""
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
# (m, n, cond, d_factor, nnz, orthogonalization, compute_Q, dtype)
CASES = [
    (10000, 200, 2.0, 2.0, 2, 0, 1, "f64"),      # test_cqrrt.cc shape family
    (2000, 100, 1e3, 1.5, 2, 0, 1, "f64"),
    (3000, 600, 100.0, 1.5, 4, 0, 1, "f64"),     # n > 256: blocked Cholesky / blocked right-solve
    (4000, 128, 50.0, 1.25, 2, 1, 1, "f64"),     # orthogonalization mode: R = R_chol
    (4000, 128, 50.0, 1.25, 2, 0, 0, "f64"),     # R-only mode: A <- A R_sk^-1
    (4000, 128, 50.0, 1.25, 4, 0, 1, "f32"),
]
for i, (m, n, cond, df, nnz, orth, cq, dt) in enumerate(CASES):
    npdt = np.float64 if dt == "f64" else np.float32
    A, st = _ref.ref_mat_gen(R, 0, m, n, n, cond, 2.0, [0] * 6, npdt)
    rc, Q, Rm, st2 = _ref.ref_cqrrt(R, A, df, st, nnz, bool(orth), bool(cq))
    out[f"ct{i}_args"] = np.array([m, n, nnz, orth, cq], dtype=np.int64)
    out[f"ct{i}_fargs"] = np.array([cond, 2.0, df])
    out[f"ct{i}_dtype"] = np.array(dt)
    out[f"ct{i}_state_in"], out[f"ct{i}_state_out"] = np.array(st, dtype=np.uint32), np.array(st2, dtype=np.uint32)
    out[f"ct{i}_rc"] = np.array([rc], dtype=np.int64)
    out[f"ct{i}_Rdiag"] = np.diag(Rm).copy()
    out[f"ct{i}_Rhead"] = np.triu(Rm)[:32, :].copy()
    out[f"ct{i}_Qhead"] = Q[:32, :].copy()
    out[f"ct{i}_Achk"] = np.array([A.sum(dtype=np.float64), np.abs(A).sum(dtype=np.float64), A[0, 0], A[-1, -1]])
    if cq and not orth:
        res = np.linalg.norm(A.astype(np.float64) - Q.astype(np.float64) @ np.triu(Rm).astype(np.float64)) / np.linalg.norm(A)
        print("cqrrt", i, (m, n), "rc", rc, "resid", res, "orth", np.linalg.norm(Q.T @ Q - np.eye(n)))
    else:
        print("cqrrt", i, (m, n), "rc", rc)
out["ct_count"] = np.array(len(CASES))
np.savez_compressed(os.path.join(HERE, "cqrrt_vectors.npz"), **out)
