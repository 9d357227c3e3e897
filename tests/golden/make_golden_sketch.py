"""Generate tests/golden/sketch_vectors.npz from the REAL reference (oracle/_ref/librl_ref.so): SparseSkOp triplets,
sparse and dense sketch-apply results.  Run in the build container only; the fixture is committed."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _ref  # noqa: E402

R = _ref.ref_lib()
assert R is not None, "build oracle/_ref first (make -C oracle ref)"
R.rlref_set_num_threads(1)
out = {}

# --- fill_sparse: (n_rows, n_cols, vec_nnz, sub, seed6)
SP = [
    (16, 100, 3, None, [0] * 6),
    (16, 100, 1, None, [5, 0, 0, 0, 7, 1]),                     # count-sketch
    (100, 16, 4, None, [0] * 6),                                 # tall operator (short axis = columns)
    (64, 1000, 8, (20, 300, 10, 200), [0xFFFFFFF0, 0xFFFFFFFF, 0, 0, 3, 4]),   # sub-block + counter carry
    (7, 7, 7, None, [0] * 6),                                    # vec_nnz == dim_major
    (4096, 20000, 1, (4096, 512, 0, 19000), [1, 2, 3, 4, 5, 6]),  # C3-shaped count-sketch, a window of columns
]
for i, (nr, nc, nnz, sub, seed) in enumerate(SP):
    rc, k, vals, rows, cols, st = _ref.ref_fill_sparse(R, nr, nc, nnz, seed, sub=sub)
    assert rc == 0
    out[f"sp{i}_args"] = np.array([nr, nc, nnz] + list(sub if sub else (nr, nc, 0, 0)), dtype=np.int64)
    out[f"sp{i}_seed"] = np.array(seed, dtype=np.uint32)
    out[f"sp{i}_rows"], out[f"sp{i}_cols"], out[f"sp{i}_vals"] = rows.astype(np.int32), cols.astype(np.int32), vals.astype(np.int8)
    out[f"sp{i}_next"] = np.array(st, dtype=np.uint32)
out["sp_count"] = np.array(len(SP))

# --- sparse left sketch on a seeded input: (S_rows, S_cols, vec_nnz, d, m, n, ro, co, alpha, beta, dtype)
AP = [
    (64, 700, 1, 64, 700, 9, 0, 0, 1.0, 0.0, "f64"),
    (64, 700, 2, 64, 700, 9, 0, 0, 1.0, 0.0, "f32"),
    (80, 1500, 4, 50, 900, 13, 7, 100, -0.5, 2.0, "f64"),
]
rng = np.random.default_rng(2024)
for i, (sr, sc, nnz, d, m, n, ro, co, alpha, beta, dt) in enumerate(AP):
    npdt = np.float64 if dt == "f64" else np.float32
    A = rng.standard_normal((m, n)).astype(npdt)
    B0 = rng.standard_normal((d, n)).astype(npdt)
    rc, B, st = _ref.ref_sketch_sparse_left(R, sr, sc, nnz, d, A, [0] * 6, alpha, beta, B0, ro, co)
    assert rc == 0
    out[f"ap{i}_args"] = np.array([sr, sc, nnz, d, m, n, ro, co], dtype=np.int64)
    out[f"ap{i}_ab"] = np.array([alpha, beta])
    out[f"ap{i}_A"], out[f"ap{i}_B0"], out[f"ap{i}_B"] = A, B0, B
    out[f"ap{i}_next"] = np.array(st, dtype=np.uint32)
out["ap_count"] = np.array(len(AP))

# --- dense sketches: (left, S_rows, S_cols, d, m, n, ro, co, family, axis, dtype)
DN = [
    (1, 24, 300, 24, 300, 7, 0, 0, 0, 0, "f64"),      # BQRRP-style wide Gaussian left sketch (generic layout)
    (1, 40, 500, 20, 260, 5, 8, 100, 1, 0, "f32"),    # uniform, sub-block
    (0, 300, 16, 16, 50, 300, 0, 0, 0, 0, "f64"),     # RS-style right sketch: (50 x 300)(300 x 16)
    (0, 16, 300, 20, 50, 12, 2, 30, 0, 1, "f64"),     # wide operator, Axis::Short, sub-block: (50 x 12)(12 x 20)
]
for i, (left, sr, sc, d, m, n, ro, co, fam, ax, dt) in enumerate(DN):
    npdt = np.float64 if dt == "f64" else np.float32
    A = rng.standard_normal((m, n)).astype(npdt)
    rc, B, st = _ref.ref_sketch_dense(R, bool(left), sr, sc, d, A, [3, 0, 0, 0, 9, 0], fam, ax, 1.0, 0.0, None, ro, co)
    assert rc == 0, _ref.ctypes.c_char_p(R.rlref_last_error())
    out[f"dn{i}_args"] = np.array([left, sr, sc, d, m, n, ro, co, fam, ax], dtype=np.int64)
    out[f"dn{i}_A"], out[f"dn{i}_B"] = A, B
    out[f"dn{i}_next"] = np.array(st, dtype=np.uint32)
out["dn_count"] = np.array(len(DN))
p = os.path.join(HERE, "sketch_vectors.npz")
np.savez_compressed(p, **out)
print("wrote", p, os.path.getsize(p), "bytes")
