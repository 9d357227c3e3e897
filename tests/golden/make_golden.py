"""Generate tests/golden/*.npz from the REAL reference (oracle/_ref/librl_ref.so, built from /root/reference).

Run in the build container only (`python tests/golden/make_golden.py`); the fixtures are committed so the
GPU box (which has no /root/reference) can check the oracle and the product against reference outputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _ref  # noqa: E402
from oracle import rl_oracle as O  # noqa: E402

R = _ref.ref_lib()
assert R is not None, "build oracle/_ref first (make -C oracle ref)"
R.rlref_set_num_threads(1)   # single-threaded BLAS => reproducible bits

out = {}
# --- fill_dense cases: (n_rows, n_cols, family, axis, layout, sub, seed6, dtype)
cases = [
    (7, 5, 0, 0, 0, None, [0, 0, 0, 0, 0, 0], "f64"),
    (5, 7, 0, 0, 0, None, [0, 0, 0, 0, 0, 0], "f64"),
    (13, 4, 1, 0, 0, None, [5, 0, 0, 0, 42, 0], "f32"),
    (64, 9, 0, 0, 0, (17, 5, 30, 3), [0xFFFFFFF0, 0xFFFFFFFF, 0, 0, 7, 1], "f64"),
    (9, 64, 0, 1, 2, (4, 33, 2, 20), [1, 2, 3, 4, 5, 6], "f32"),
    (256, 32, 0, 0, 0, None, [0, 0, 0, 0, 0, 0], "f64"),       # RS Omega of config C1 (n=256, k=32)
    (100, 3, 0, 0, 1, None, [9, 0, 0, 0, 1, 0], "f64"),
]
for i, (nr, nc, fam, ax, lay, sub, seed, dt) in enumerate(cases):
    npdt = np.float64 if dt == "f64" else np.float32
    rc, buf, nxt = _ref.fill_dense(R, "rlref", nr, nc, seed, npdt, fam, ax, lay, sub)
    assert rc == 0
    out[f"fill{i}_args"] = np.array([nr, nc, fam, ax, lay] + list(sub if sub else (nr, nc, 0, 0)), dtype=np.int64)
    out[f"fill{i}_seed"] = np.array(seed, dtype=np.uint32)
    out[f"fill{i}_buf"] = buf
    out[f"fill{i}_next"] = np.array(nxt, dtype=np.uint32)
out["fill_count"] = np.array(len(cases))

# --- RSVD on the reference's own kind of test input (polynomial decay, test_rf.cc:159-162 / test_rsvd.cc:181-184)
rs_cases = [
    # m, n, k, p, q, block, stab, cond, exponent
    (10, 10, 5, 10, 1, 2, 0, 2.0, 1.0),          # test/drivers/test_rsvd.cc:168-197 (SimpleTest)
    (400, 64, 16, 0, 1, 16, 0, 2025.0, 2.0),
    (400, 64, 16, 2, 1, 16, 1, 2025.0, 2.0),
    (400, 64, 16, 3, 1, 8, 1, 2025.0, 2.0),
    (4096, 256, 32, 0, 1, 32, 1, 2025.0, 2.0),   # BASELINE.json configs[0]
]
for i, (m, n, k, p, q, b, stab, cond, expo) in enumerate(rs_cases):
    A, st1 = _ref.ref_mat_gen(R, 0, m, n, n if m > 10 else k, cond, expo, [0] * 6)
    o = O.StackOpts(p, q, b, stab, 1, 1, False, False)
    rc, kk, U, S, V, st2 = _ref.ref_rsvd(R, A, k, 0.0, st1, o)
    out[f"rsvd{i}_args"] = np.array([m, n, k, p, q, b, stab], dtype=np.int64)
    out[f"rsvd{i}_cond_expo"] = np.array([cond, expo])
    out[f"rsvd{i}_state_in"] = np.array(st1, dtype=np.uint32)
    out[f"rsvd{i}_state_out"] = np.array(st2, dtype=np.uint32)
    out[f"rsvd{i}_rc_k"] = np.array([rc, kk], dtype=np.int64)
    out[f"rsvd{i}_S"] = S
    out[f"rsvd{i}_V"] = V
    out[f"rsvd{i}_Uhead"] = U[:64].copy()
    out[f"rsvd{i}_resid"] = np.array([np.linalg.norm(A - (U * S) @ V.T) / np.linalg.norm(A)])
    out[f"rsvd{i}_Achk"] = np.array([A.sum(), np.abs(A).sum(), A[0, 0], A[-1, -1]])
out["rsvd_count"] = np.array(len(rs_cases))
np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
print("wrote", os.path.join(HERE, "reference_vectors.npz"), os.path.getsize(os.path.join(HERE, "reference_vectors.npz")), "bytes")
