"""Generate tests/golden/skgen_vectors.npz from the REAL reference (oracle/_ref/librl_ref.so): RandBLAS::sketch_general with a DenseSkOp for
every layout / opS / opA combination, left (skge.hh:859-905) and right (:1031-1076), with submatrix offsets, padded leading dimensions and
alpha / beta != (1, 0).  Inputs come from numpy's frozen RandomState stream (seed 1000 + case), only the outputs are stored.  Run in the build container only."""
import itertools
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
i = 0
for left, layout, opS, opA in itertools.product((1, 0), (1, 2), (0, 1), (0, 1)):
    for (d, n, m, ro, co, dt, fam, ax) in ((24, 37, 150, 3, 5, np.float64, 0, 0), (8, 5, 33, 0, 2, np.float32, 1, 1)):
        rs, cs = ((m, d) if opS else (d, m)) if left else ((d, n) if opS else (n, d))
        S_rows, S_cols = rs + ro + 2, cs + co + 1
        ra, ca = (n, m) if opA else (m, n)
        rb, cb = (d, n) if left else (m, d)
        lda = (ra if layout == 1 else ca) + 3
        ldb = (rb if layout == 1 else cb) + 2
        rng = np.random.RandomState(1000 + i)
        A = rng.standard_normal(lda * (ca if layout == 1 else ra)).astype(dt)
        B = rng.standard_normal(ldb * (cb if layout == 1 else rb)).astype(dt)
        seed = [7 + i, 0, 0, 0, 11, 0]
        rc, Bo, st = _ref.ref_sketch_general_dense(R, left, layout, opS, opA, (S_rows, S_cols, fam, ax), (d, n, m), A, lda, B, ldb, seed, alpha=0.75,
                                                   beta=-0.5, ro=ro, co=co)
        assert rc == 0
        out[f"sg{i}_args"] = np.array([left, layout, opS, opA, d, n, m, ro, co, S_rows, S_cols, fam, ax, lda, ldb], dtype=np.int64)
        out[f"sg{i}_Bout"] = Bo
        out[f"sg{i}_dtype"] = np.array("f64" if dt == np.float64 else "f32")
        out[f"sg{i}_state_in"], out[f"sg{i}_state_out"] = np.array(seed, dtype=np.uint32), np.array(st, dtype=np.uint32)
        i += 1
out["sg_count"] = np.array(i)
# ---- the same with a short-axis SparseSkOp (skge.hh:907-960, 1078-1131): op(submat(S)) wide, i.e. S tall whenever its transpose is applied
j = 0
for left, layout, opS, opA in itertools.product((1, 0), (1, 2), (0, 1), (0, 1)):
    for (d, n, m, ro, co, dt, nnz) in ((24, 37, 150, 3, 5, np.float64, 3), (8, 5, 33, 0, 2, np.float32, 1)):
        if not left:
            n = 60 if d == 24 else 21         # right sketch: contraction length n > d so that the n x d block is tall
        rs, cs = ((m, d) if opS else (d, m)) if left else ((d, n) if opS else (n, d))
        S_rows, S_cols = rs + ro + 2, cs + co + 1
        assert (S_rows > S_cols) == (rs > cs)
        ra, ca = (n, m) if opA else (m, n)
        rb, cb = (d, n) if left else (m, d)
        lda = (ra if layout == 1 else ca) + 3
        ldb = (rb if layout == 1 else cb) + 2
        rng = np.random.RandomState(2000 + j)
        A = rng.standard_normal(lda * (ca if layout == 1 else ra)).astype(dt)
        B = rng.standard_normal(ldb * (cb if layout == 1 else rb)).astype(dt)
        seed = [5 + j, 0, 0, 0, 13, 0]
        rc, Bo, st = _ref.ref_sketch_general_sparse(R, left, layout, opS, opA, (S_rows, S_cols, nnz), (d, n, m), A, lda, B, ldb, seed, alpha=0.75,
                                                    beta=-0.5, ro=ro, co=co)
        assert rc == 0
        out[f"ss{j}_args"] = np.array([left, layout, opS, opA, d, n, m, ro, co, S_rows, S_cols, nnz, lda, ldb], dtype=np.int64)
        out[f"ss{j}_Bout"] = Bo
        out[f"ss{j}_dtype"] = np.array("f64" if dt == np.float64 else "f32")
        out[f"ss{j}_state_in"], out[f"ss{j}_state_out"] = np.array(seed, dtype=np.uint32), np.array(st, dtype=np.uint32)
        j += 1
out["ss_count"] = np.array(j)
# ---- Axis::Long operators (LASO, sparse_skops.hh:669-704): COO export (wide / tall / square, sub-matrices, many duplicates, fp32) and the left sketch
LA = [(20, 300, 5, np.float64, None), (300, 20, 4, np.float64, None), (16, 40, 30, np.float32, None), (20, 300, 6, np.float64, (7, 100, 3, 50)),
      (50, 50, 8, np.float64, (20, 30, 5, 10)), (64, 5000, 64, np.float64, None)]
for k, (r, c, nnz, dt, sub) in enumerate(LA):
    seed = [3 + k, 0, 0, 0, 17, 0]
    rc, nz, vals, rows, cols, st = _ref.ref_fill_sparse(R, r, c, nnz, seed, dt, axis=0, sub=sub)
    assert rc == 0
    out[f"la{k}_args"] = np.array([r, c, nnz] + list(sub if sub else (r, c, 0, 0)), dtype=np.int64)
    out[f"la{k}_dtype"] = np.array("f64" if dt == np.float64 else "f32")
    out[f"la{k}_vals"], out[f"la{k}_rows"], out[f"la{k}_cols"] = vals[:nz], rows[:nz], cols[:nz]
    out[f"la{k}_state_in"], out[f"la{k}_state_out"] = np.array(seed, dtype=np.uint32), np.array(st, dtype=np.uint32)
out["la_count"] = np.array(len(LA))
LS = [(24, 37, 150, 3, 5, np.float64, 6), (8, 5, 33, 0, 2, np.float32, 20), (128, 64, 3000, 1, 7, np.float64, 8)]
for k, (d, n, m, ro, co, dt, nnz) in enumerate(LS):
    S_rows, S_cols = d + ro + 2, m + co + 1
    lda, ldb = m + 3, d + 2
    rng = np.random.RandomState(3000 + k)
    A = rng.standard_normal(lda * n).astype(dt)
    B = rng.standard_normal(ldb * n).astype(dt)
    seed = [9 + k, 0, 0, 0, 23, 0]
    rc, Bo, st = _ref.ref_sketch_general_sparse(R, 1, 1, 0, 0, (S_rows, S_cols, nnz, 0), (d, n, m), A, lda, B, ldb, seed, alpha=0.75, beta=-0.5, ro=ro, co=co)
    assert rc == 0
    out[f"ls{k}_args"] = np.array([d, n, m, ro, co, S_rows, S_cols, nnz, lda, ldb], dtype=np.int64)
    out[f"ls{k}_dtype"] = np.array("f64" if dt == np.float64 else "f32")
    out[f"ls{k}_Bout"] = Bo
    out[f"ls{k}_state_in"], out[f"ls{k}_state_out"] = np.array(seed, dtype=np.uint32), np.array(st, dtype=np.uint32)
out["ls_count"] = np.array(len(LS))
np.savez_compressed(os.path.join(HERE, "skgen_vectors.npz"), **out)
print("cases", i, j)
