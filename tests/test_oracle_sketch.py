"""CPU: the oracle's restatement of RandBLAS's sparse/dense sketching operators against (1) the golden vectors produced by the
real reference (tests/golden/sketch_vectors.npz, made by tests/golden/make_golden_sketch.py) and (2) the compiled reference
itself when oracle/_ref exists.  Integer work (triplets, RNG state) is bit-exact; applied sketches are fp sums whose order differs
(the reference reassociates with OpenMP simd), so they are compared to a few ulps of the accumulated magnitude."""
import os

import numpy as np
import pytest

import _ref
from oracle import rl_oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sketch_vectors.npz"))


def _state(seed):
    return O.RNGState((int(seed[4]), int(seed[5])), [int(x) for x in seed[:4]])


@pytest.mark.parametrize("i", range(int(G["sp_count"])))
def test_fill_sparse_golden(i):
    nr, nc, nnz, sr, sc, ro, co = [int(x) for x in G[f"sp{i}_args"]]
    k, vals, rows, cols, nxt = O.fill_sparse(nr, nc, nnz, _state(G[f"sp{i}_seed"]), sub=(sr, sc, ro, co))
    assert k == len(G[f"sp{i}_rows"])
    assert np.array_equal(rows, G[f"sp{i}_rows"]) and np.array_equal(cols, G[f"sp{i}_cols"]) and np.array_equal(vals, G[f"sp{i}_vals"])
    assert list(nxt.words()) == list(G[f"sp{i}_next"])


def test_saso_structure():
    """RandBLAS/test/datastructures/test_sparseskop.cc:60-95: every long-axis vector has exactly vec_nnz distinct short-axis
    indices, values are +-1."""
    for (nr, nc, nnz) in [(16, 100, 3), (100, 16, 4), (50, 50, 50), (4096, 3000, 1)]:
        k, vals, rows, cols, _ = O.fill_sparse(nr, nc, nnz, O.RNGState(1))
        assert k == nnz * max(nr, nc) and set(np.unique(vals)) <= {-1.0, 1.0}
        major, minor = (rows, cols) if nr <= nc else (cols, rows)
        assert major.min() >= 0 and major.max() < min(nr, nc)
        for v in range(0, max(nr, nc), 7):
            blk = major[minor == v]
            assert len(blk) == nnz and len(set(blk)) == nnz and np.all(np.diff(blk) > 0)


@pytest.mark.parametrize("i", range(int(G["ap_count"])))
def test_sparse_apply_golden(i):
    sr, sc, nnz, d, m, n, ro, co = [int(x) for x in G[f"ap{i}_args"]]
    alpha, beta = G[f"ap{i}_ab"]
    A, B0, Bref = G[f"ap{i}_A"], G[f"ap{i}_B0"], G[f"ap{i}_B"]
    B, nxt = O.sketch_sparse_left(sr, sc, nnz, d, A, O.RNGState(0), alpha, beta, B0, ro, co)
    tol = 50 * np.finfo(A.dtype).eps * (np.abs(A).max() * nnz * m / d + np.abs(B0).max() * abs(beta))
    assert np.abs(B - Bref).max() <= tol
    assert list(nxt.words()) == list(G[f"ap{i}_next"])


@pytest.mark.parametrize("i", range(int(G["dn_count"])))
def test_dense_apply_golden(i):
    left, sr, sc, d, m, n, ro, co, fam, ax = [int(x) for x in G[f"dn{i}_args"]]
    A, Bref = G[f"dn{i}_A"], G[f"dn{i}_B"]
    st = O.RNGState((9, 0), (3, 0, 0, 0))
    if left:
        B, nxt = O.sketch_dense_left(sr, sc, d, A, st, fam, ax, ro=ro, co=co)
    else:
        B, nxt = O.sketch_dense_right(A, sr, sc, d, st, fam, ax, ro=ro, co=co)
    assert np.abs(B - Bref).max() <= 100 * np.finfo(A.dtype).eps * np.abs(Bref).max()
    assert list(nxt.words()) == list(G[f"dn{i}_next"])


@pytest.mark.skipif(_ref.ref_lib() is None, reason="compiled reference not present")
def test_fill_sparse_vs_compiled_reference():
    L = _ref.ref_lib()
    rng = np.random.default_rng(5)
    for _ in range(25):
        nr, nc = int(rng.integers(1, 60)), int(rng.integers(1, 400))
        if rng.random() < 0.3:
            nr, nc = nc, nr
        nnz = int(rng.integers(1, min(nr, nc, 9) + 1))
        sr, sc = int(rng.integers(1, nr + 1)), int(rng.integers(1, nc + 1))
        sub = (sr, sc, int(rng.integers(0, nr - sr + 1)), int(rng.integers(0, nc - sc + 1)))
        seed = [int(x) for x in rng.integers(0, 2 ** 32, 6)]
        rc, k, vals, rows, cols, st = _ref.ref_fill_sparse(L, nr, nc, nnz, seed, sub=sub)
        k2, v2, r2, c2, st2 = O.fill_sparse(nr, nc, nnz, _state(seed), sub=sub)
        assert rc == 0 and k == k2 and np.array_equal(rows, r2) and np.array_equal(cols, c2) and np.array_equal(vals, v2)
        assert st == list(st2.words())


from _qrcases import GS, sg_input  # noqa: E402


@pytest.mark.parametrize("i", range(int(GS["sg_count"])))
def test_sketch_general_dense_all_flags_golden(i):
    """sketch_general with a DenseSkOp for every layout / opS / opA, left and right (skge.hh:859-905, 1031-1076): the restatement against
    golden outputs of the compiled reference (padded leading dimensions, submatrix offsets, alpha = 0.75, beta = -0.5; the padding of B
    must come back untouched).  fp64 1e-12, fp32 2e-5 (relative to the largest entry)."""
    c = sg_input(i)
    out, nxt = O.sketch_general_dense(c["left"], c["layout"], c["opS"], c["opA"], c["d"], c["n"], c["m"], 0.75, c["S_rows"], c["S_cols"], c["ro"],
                                      c["co"], c["A"], c["lda"], -0.5, c["B"], c["ldb"], O.RNGState.from_words(c["seed"]), c["family"], c["axis"])
    ref = GS[f"sg{i}_Bout"]
    assert list(nxt.words()) == list(GS[f"sg{i}_state_out"])
    tol = 1e-12 if c["dtype"] == np.float64 else 2e-5
    assert np.abs(out - ref).max() <= tol * np.abs(ref).max()


from _qrcases import ss_input  # noqa: E402


@pytest.mark.parametrize("i", range(int(GS["ss_count"])))
def test_sketch_general_sparse_all_flags_golden(i):
    """The same with a short-axis SparseSkOp (skge.hh:907-960, 1078-1131): restatement vs golden outputs of the compiled reference.
    This also pins the fact the device path relies on: a tall short-axis operator is the transpose of the wide one with the same seed."""
    c = ss_input(i)
    out, nxt = O.sketch_general_sparse(c["left"], c["layout"], c["opS"], c["opA"], c["d"], c["n"], c["m"], 0.75, c["S_rows"], c["S_cols"],
                                       c["vec_nnz"], c["ro"], c["co"], c["A"], c["lda"], -0.5, c["B"], c["ldb"], O.RNGState.from_words(c["seed"]))
    ref = GS[f"ss{i}_Bout"]
    assert list(nxt.words()) == list(GS[f"ss{i}_state_out"])
    tol = 1e-12 if c["dtype"] == np.float64 else 2e-5
    assert np.abs(out - ref).max() <= tol * np.abs(ref).max()


from _qrcases import ls_input  # noqa: E402


@pytest.mark.parametrize("k", range(int(GS["la_count"])))
def test_fill_sparse_laso_golden(k):
    """Axis::Long operators (sparse_skops.hh:669-704): the COO triplets of the restatement are bit-exact with the compiled reference's
    (indices, merged sqrt(count) values in the working precision, order, returned state)."""
    r, c, nnz, sr, sc, ro, co = [int(x) for x in GS[f"la{k}_args"]]
    dt = np.float64 if str(GS[f"la{k}_dtype"]) == "f64" else np.float32
    nz, vals, rows, cols, st = O.fill_sparse_laso(r, c, nnz, O.RNGState.from_words([int(x) for x in GS[f"la{k}_state_in"]]), dt, sub=(sr, sc, ro, co))
    assert nz == len(GS[f"la{k}_vals"]) and list(st.words()) == list(GS[f"la{k}_state_out"])
    assert np.array_equal(vals, GS[f"la{k}_vals"]) and np.array_equal(rows, GS[f"la{k}_rows"]) and np.array_equal(cols, GS[f"la{k}_cols"])


@pytest.mark.parametrize("k", range(int(GS["ls_count"])))
def test_sketch_sparse_left_laso_golden(k):
    c = ls_input(k)
    out, nxt = O.sketch_general_sparse(1, O.LAYOUT_COLMAJOR, 0, 0, c["d"], c["n"], c["m"], 0.75, c["S_rows"], c["S_cols"], c["vec_nnz"], c["ro"], c["co"],
                                       c["A"], c["lda"], -0.5, c["B"], c["ldb"], O.RNGState.from_words(c["seed"]), major_axis=O.AXIS_LONG)
    ref = GS[f"ls{k}_Bout"]
    assert list(nxt.words()) == list(GS[f"ls{k}_state_out"])
    tol = 1e-12 if c["dtype"] == np.float64 else 2e-5
    assert np.abs(out - ref).max() <= tol * np.abs(ref).max()
