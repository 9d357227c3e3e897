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
np.savez_compressed(os.path.join(HERE, "skgen_vectors.npz"), **out)
print("cases", i, j)
