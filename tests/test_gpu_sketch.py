"""GPU parity, RandBLAS sketch-apply through the C-ABI.

Tolerances (stated):
 * SparseSkOp triplets (rows, cols, +-1 values) and RNG-state advancement are integer work: bit-exact against the oracle and
   against the reference-generated golden vectors.
 * Applied sparse sketch: sums of +-A entries; the device adds in ascending source-row order, the reference in a reassociated
   (OpenMP simd) order, so B is compared to 50 eps * (largest magnitude a sum of its length can reach).
 * Dense sketch: the device's Gaussian entries are within a few float ulps of the host libm's (tests/test_gpu_fill.py), so against
   the oracle/golden B is compared to 2e-6 relative (Frobenius); against a product with the DEVICE-generated operator to 1e-12."""
import os

import numpy as np
import pytest
import torch

import randlapack_b200 as rl
from oracle import rl_oracle as O

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sketch_vectors.npz"))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a.T)).cuda().t()


def host(t):
    return np.asfortranarray(t.cpu().numpy())


def _st(seed):
    return rl.RNGState(key=(int(seed[4]), int(seed[5])), counter=[int(x) for x in seed[:4]])


def _ost(seed):
    return O.RNGState((int(seed[4]), int(seed[5])), [int(x) for x in seed[:4]])


@pytest.mark.parametrize("i", range(int(G["sp_count"])))
def test_fill_sparse_golden(ctx, i):
    nr, nc, nnz, sr, sc, ro, co = [int(x) for x in G[f"sp{i}_args"]]
    k, vals, rows, cols, nxt = rl.fill_sparse(ctx, rl.SparseDist(nr, nc, nnz), _st(G[f"sp{i}_seed"]), torch.float64, (sr, sc, ro, co))
    assert k == len(G[f"sp{i}_rows"])
    assert np.array_equal(rows.cpu().numpy(), G[f"sp{i}_rows"]) and np.array_equal(cols.cpu().numpy(), G[f"sp{i}_cols"])
    assert np.array_equal(vals.cpu().numpy(), G[f"sp{i}_vals"])
    assert list(nxt.words()) == list(G[f"sp{i}_next"])


def test_fill_sparse_vs_oracle_random(ctx):
    rng = np.random.default_rng(17)
    for trial in range(30):
        nr, nc = int(rng.integers(1, 70)), int(rng.integers(1, 3000))
        if rng.random() < 0.3:
            nr, nc = nc, nr
        nnz = int(rng.integers(1, min(nr, nc, 12) + 1))
        sr, sc = int(rng.integers(1, nr + 1)), int(rng.integers(1, nc + 1))
        sub = (sr, sc, int(rng.integers(0, nr - sr + 1)), int(rng.integers(0, nc - sc + 1)))
        seed = [int(x) for x in rng.integers(0, 2 ** 32, 6)]
        if trial % 3 == 0:
            seed[0], seed[1] = 0xFFFFFFF0, 0xFFFFFFFF
        dt = torch.float32 if trial % 2 else torch.float64
        k, vals, rows, cols, nxt = rl.fill_sparse(ctx, rl.SparseDist(nr, nc, nnz), _st(seed), dt, sub)
        k2, v2, r2, c2, st2 = O.fill_sparse(nr, nc, nnz, _ost(seed), sub=sub)
        assert k == k2, (nr, nc, nnz, sub)
        assert np.array_equal(rows.cpu().numpy(), r2) and np.array_equal(cols.cpu().numpy(), c2)
        assert np.array_equal(vals.cpu().numpy().astype(np.float64), v2)
        assert list(nxt.words()) == list(st2.words())


@pytest.mark.parametrize("i", range(int(G["ap_count"])))
def test_sparse_apply_golden(ctx, i):
    sr, sc, nnz, d, m, n, ro, co = [int(x) for x in G[f"ap{i}_args"]]
    alpha, beta = [float(x) for x in G[f"ap{i}_ab"]]
    A, B0, Bref = G[f"ap{i}_A"], G[f"ap{i}_B0"], G[f"ap{i}_B"]
    st = rl.RNGState(0)
    B = rl.sketch_general_left(ctx, rl.SparseDist(sr, sc, nnz), st, dev(A), d, alpha, beta, dev(B0), ro, co)
    tol = 50 * np.finfo(A.dtype).eps * (np.abs(A).max() * nnz * m / d + np.abs(B0).max() * abs(beta)) * max(1.0, abs(alpha))
    assert np.abs(host(B) - Bref).max() <= tol
    assert list(st.words()) == list(G[f"ap{i}_next"])


# (d, m, n, nnz): covers every rows-per-thread variant of the kernel (d <= 512, 1024, 4096, 16384), ragged m (not a multiple of
# the 2048-row chunk), n not a multiple of the column tile, single-row / single-column inputs
APPLY = [(1, 1, 1, 1), (16, 16, 3, 16), (16, 333, 5, 3), (600, 5000, 17, 2), (2000, 9000, 9, 1), (4096, 20000, 24, 1), (5000, 12000, 6, 4),
         (64, 70000, 33, 8), (300, 4097, 8, 16),
         # strip kernel (16-byte aligned A, whole 512/256-row chunks + ragged tail): 1 / 2 / 4 / 8 row blocks per cluster, partial strips
         (4096, 65536, 70, 4), (1024, 33000, 40, 1), (3000, 16384, 32, 2), (8192, 40960, 31, 1)]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("case", APPLY)
def test_sparse_apply_vs_oracle(ctx, dtype, case):
    d, m, n, nnz = case
    npdt = np.float64 if dtype == torch.float64 else np.float32
    rng = np.random.default_rng(d + m)
    A = np.asfortranarray(rng.standard_normal((m, n)).astype(npdt))
    st, ost = rl.RNGState(3), O.RNGState(3)
    B = rl.sketch_general_left(ctx, rl.SparseDist(d, m, nnz), st, dev(A))
    Bo, onxt = O.sketch_sparse_left(d, m, nnz, d, A, ost)
    tol = 50 * np.finfo(npdt).eps * np.abs(A).max() * max(4.0, 8.0 * nnz * m / d)
    assert np.abs(host(B) - Bo).max() <= tol, np.abs(host(B) - Bo).max()
    assert list(st.words()) == list(onxt.words())
    # run-to-run determinism (fixed summation order)
    B2 = rl.sketch_general_left(ctx, rl.SparseDist(d, m, nnz), rl.RNGState(3), dev(A))
    assert torch.equal(B, B2)


def test_sparse_apply_properties_large(ctx):
    """Size-independent properties at a size the oracle does not run at: linearity, and the checksum
    1^T (S A) = (1^T S) A with 1^T S taken from the device's own (bit-exact-tested) triplets."""
    d, m, n, nnz = 4096, 1 << 20, 64, 2
    g = torch.Generator(device="cuda").manual_seed(1)
    A1 = rl.to_f(torch.randn((m, n), dtype=torch.float32, device="cuda", generator=g))
    A2 = rl.to_f(torch.randn((m, n), dtype=torch.float32, device="cuda", generator=g))
    D = rl.SparseDist(d, m, nnz)
    B1 = rl.sketch_general_left(ctx, D, rl.RNGState(9), A1)
    B2 = rl.sketch_general_left(ctx, D, rl.RNGState(9), A2)
    B12 = rl.sketch_general_left(ctx, D, rl.RNGState(9), rl.to_f(A1 + A2))
    scale = float(nnz * m / d) ** 0.5
    assert (B12 - (B1 + B2)).abs().max().item() <= 1e-5 * scale * 8
    k, vals, rows, cols, _ = rl.fill_sparse(ctx, D, rl.RNGState(9), torch.float32)
    colsum = torch.zeros(m, dtype=torch.float64, device="cuda").index_add_(0, cols, vals.double())
    lhs = B1.double().sum(dim=0)
    rhs = colsum @ A1.double()
    assert (lhs - rhs).abs().max().item() <= 1e-3 * (nnz * m) ** 0.5


@pytest.mark.parametrize("i", range(int(G["dn_count"])))
def test_dense_apply_golden(ctx, i):
    left, sr, sc, d, m, n, ro, co, fam, ax = [int(x) for x in G[f"dn{i}_args"]]
    A, Bref = G[f"dn{i}_A"], G[f"dn{i}_B"]
    st = rl.RNGState(key=(9, 0), counter=(3, 0, 0, 0))
    D = rl.DenseDist(sr, sc, fam, ax)
    if left:
        B = rl.sketch_general_left(ctx, D, st, dev(A), d, ro_s=ro, co_s=co)
    else:
        B = rl.sketch_general_right(ctx, dev(A), D, st, d, ro_s=ro, co_s=co)
    tol = 2e-6 if A.dtype == np.float64 else 2e-5
    assert np.linalg.norm(host(B) - Bref) <= tol * np.linalg.norm(Bref)
    assert list(st.words()) == list(G[f"dn{i}_next"])


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_dense_left_never_materialised_matches_materialised(ctx, dtype):
    """S (d x m) would be 2.1 GB here in fp64; the kernel regenerates it in 32 MB panels.  Compare against the product with
    explicitly materialised row-blocks of the same operator, and check alpha/beta handling."""
    d, m, n = 128, 1 << 21, 32
    if dtype == torch.float32:
        d, m = 256, 1 << 19
    g = torch.Generator(device="cuda").manual_seed(4)
    A = rl.to_f(torch.randn((m, n), dtype=dtype, device="cuda", generator=g))
    B0 = rl.to_f(torch.randn((d, n), dtype=dtype, device="cuda", generator=g))
    D = rl.DenseDist(d, m)
    st = rl.RNGState(21)
    B = rl.sketch_general_left(ctx, D, st, A, d, alpha=0.5, beta=-1.0, B=B0.clone())
    exp = -1.0 * B0.double()
    step = 1 << 17
    for j0 in range(0, m, step):
        buf, _ = rl.fill_dense(ctx, D, rl.RNGState(21), dtype, rl.LAYOUT_COLMAJOR, (d, step, 0, j0))
        S = buf.view(step, d).t()
        exp += 0.5 * (S.double() @ A[j0:j0 + step].double())
    rel = ((B.double() - exp).norm() / exp.norm()).item()
    assert rel <= (1e-12 if dtype == torch.float64 else 1e-5), rel
    _, nxt = rl.fill_dense(ctx, D, rl.RNGState(21), dtype, rl.LAYOUT_NATURAL, (1, 1, 0, 0))
    assert st == O_next(d, m, 21)


def O_next(nr, nc, key):
    s = O.dense_next_state(nr, nc, O.AXIS_LONG, O.RNGState(key))
    return rl.RNGState(s.key, s.counter)


# ---- sketch_general with every layout / transposition flag (skge.hh:859-905 left, 1031-1076 right) -----------------------------------------
from _qrcases import GS, sg_input  # noqa: E402


def _sg_run(ctx, c, A, B, alpha, beta, seed):
    Ad, Bd = torch.from_numpy(A).cuda(), torch.from_numpy(B.copy()).cuda()
    D = rl.DenseDist(c["S_rows"], c["S_cols"], c["family"], c["axis"])
    s = _st(seed)
    if c["left"]:
        rl.sketch_general_dense_left(ctx, c["layout"], c["opS"], c["opA"], c["d"], c["n"], c["m"], alpha, D, c["ro"], c["co"], Ad, c["lda"], beta,
                                     Bd, c["ldb"], s)
    else:
        rl.sketch_general_dense_right(ctx, c["layout"], c["opA"], c["opS"], c["m"], c["d"], c["n"], alpha, Ad, c["lda"], D, c["ro"], c["co"], beta,
                                      Bd, c["ldb"], s)
    return Bd.cpu().numpy(), s


@pytest.mark.parametrize("i", range(int(GS["sg_count"])))
def test_sketch_general_dense_all_flags_golden(ctx, i):
    """Every layout / opS / opA combination, left and right, against golden outputs of the compiled reference (padded leading dimensions,
    submatrix offsets, alpha = 0.75, beta = -0.5).  RNG state exact; the padding of B untouched (bit-exact); values 2e-6 (fp64) / 2e-5 (fp32)
    relative to the largest entry (the device's Gaussians are within a few float ulps of the host libm's)."""
    c = sg_input(i)
    out, s = _sg_run(ctx, c, c["A"], c["B"], 0.75, -0.5, c["seed"])
    ref = GS[f"sg{i}_Bout"]
    assert list(s.words()) == list(GS[f"sg{i}_state_out"])
    rb, cb = (c["d"], c["n"]) if c["left"] else (c["m"], c["d"])
    mask = np.ones(ref.shape, dtype=bool)
    O._mat_view(mask, rb, cb, c["ldb"], c["layout"])[:, :] = False          # True on the padding
    assert np.array_equal(out[mask], c["B"][mask]), "padding of B was written"
    tol = 2e-6 if c["dtype"] == np.float64 else 2e-5
    assert np.abs(out - ref).max() <= tol * np.abs(ref).max()


@pytest.mark.parametrize("flags", [(1, 2, 0, 0), (1, 1, 1, 1), (1, 2, 1, 1), (0, 2, 0, 0), (0, 1, 1, 1)])
def test_sketch_general_dense_tall_vs_oracle(ctx, flags):
    """The same at a size where the left sketch runs on the digit-slice engine (m = 20000 rows, d = 64, n = 96) and the data matrix has to be
    transposed / the result written transposed; against the restatement on the same operator (2e-6 relative, Frobenius)."""
    left, layout, opS, opA = flags
    d, n, m = 64, 96, 20000
    rs, cs = ((m, d) if opS else (d, m)) if left else ((d, n) if opS else (n, d))
    ra, ca = (n, m) if opA else (m, n)
    rb, cb = (d, n) if left else (m, d)
    c = dict(left=left, layout=layout, opS=opS, opA=opA, d=d, n=n, m=m, ro=1, co=2, S_rows=rs + 1, S_cols=cs + 2, family=0, axis=0,
             lda=(ra if layout == 1 else ca) + 1, ldb=(rb if layout == 1 else cb))
    rng = np.random.RandomState(5)
    A = rng.standard_normal(c["lda"] * (ca if layout == 1 else ra))
    B = rng.standard_normal(c["ldb"] * (cb if layout == 1 else rb))
    seed = [3, 0, 0, 0, 9, 0]
    out, s = _sg_run(ctx, c, A, B, 1.0, 0.5, seed)
    ref, nxt = O.sketch_general_dense(left, layout, opS, opA, d, n, m, 1.0, c["S_rows"], c["S_cols"], 1, 2, A, c["lda"], 0.5, B, c["ldb"],
                                      O.RNGState.from_words(seed))
    assert list(s.words()) == list(nxt.words())
    assert np.linalg.norm(out - ref) <= 2e-6 * np.linalg.norm(ref)


from _qrcases import ss_input  # noqa: E402


def _ss_run(ctx, c, A, B, alpha, beta, seed):
    Ad, Bd = torch.from_numpy(A).cuda(), torch.from_numpy(B.copy()).cuda()
    D = rl.SparseDist(c["S_rows"], c["S_cols"], c["vec_nnz"])
    s = _st(seed)
    if c["left"]:
        rl.sketch_general_sparse_left(ctx, c["layout"], c["opS"], c["opA"], c["d"], c["n"], c["m"], alpha, D, c["ro"], c["co"], Ad, c["lda"], beta,
                                      Bd, c["ldb"], s)
    else:
        rl.sketch_general_sparse_right(ctx, c["layout"], c["opA"], c["opS"], c["m"], c["d"], c["n"], alpha, Ad, c["lda"], D, c["ro"], c["co"], beta,
                                       Bd, c["ldb"], s)
    return Bd.cpu().numpy(), s


@pytest.mark.parametrize("i", range(int(GS["ss_count"])))
def test_sketch_general_sparse_all_flags_golden(ctx, i):
    """Short-axis SparseSkOp, every layout / opS / opA, left (skge.hh:907-960) and right (:1078-1131; run as the left sketch of the transposed
    problem), against golden outputs of the compiled reference.  RNG state exact, padding of B untouched, values 1e-12 (fp64) / 2e-5 (fp32)
    relative to the largest entry (sums of at most a few dozen +-A entries in a different order)."""
    c = ss_input(i)
    out, s = _ss_run(ctx, c, c["A"], c["B"], 0.75, -0.5, c["seed"])
    ref = GS[f"ss{i}_Bout"]
    assert list(s.words()) == list(GS[f"ss{i}_state_out"])
    rb, cb = (c["d"], c["n"]) if c["left"] else (c["m"], c["d"])
    mask = np.ones(ref.shape, dtype=bool)
    O._mat_view(mask, rb, cb, c["ldb"], c["layout"])[:, :] = False
    assert np.array_equal(out[mask], c["B"][mask]), "padding of B was written"
    tol = 1e-12 if c["dtype"] == np.float64 else 2e-5
    assert np.abs(out - ref).max() <= tol * np.abs(ref).max()


def test_sketch_general_sparse_rejects_tall_op(ctx):
    """op(submat(S)) tall (a wide S transposed on the left) is not a sketch: RLB200_ERR_UNSUPPORTED, never a fallback."""
    A = torch.zeros(40 * 8, dtype=torch.float64, device="cuda")
    B = torch.zeros(40 * 8, dtype=torch.float64, device="cuda")
    with pytest.raises(rl.Error):
        rl.sketch_general_sparse_left(ctx, rl.LAYOUT_COLMAJOR, True, False, 20, 8, 10, 1.0, rl.SparseDist(10, 40, 2), 0, 0, A, 10, 0.0, B, 20,
                                      rl.RNGState(0))


@pytest.mark.parametrize("flags", [(1, 2, 0, 0), (1, 1, 1, 1), (0, 1, 0, 0), (0, 2, 1, 1)])
def test_sketch_general_sparse_tall_vs_oracle(ctx, flags):
    """The same at a size where the strip kernel runs (20000-row data matrix, d = 256, 128 columns, vec_nnz = 2) behind the transposition
    wrappers, against the restatement on the same operator."""
    left, layout, opS, opA = flags
    d, n, m = (256, 128, 20000) if left else (64, 2000, 300)
    rs, cs = ((m, d) if opS else (d, m)) if left else ((d, n) if opS else (n, d))
    ra, ca = (n, m) if opA else (m, n)
    rb, cb = (d, n) if left else (m, d)
    c = dict(left=left, layout=layout, opS=opS, opA=opA, d=d, n=n, m=m, ro=1, co=2, S_rows=rs + 1, S_cols=cs + 2, vec_nnz=2,
             lda=(ra if layout == 1 else ca), ldb=(rb if layout == 1 else cb) + 1)
    rng = np.random.RandomState(6)
    A = rng.standard_normal(c["lda"] * (ca if layout == 1 else ra))
    B = rng.standard_normal(c["ldb"] * (cb if layout == 1 else rb))
    seed = [4, 0, 0, 0, 21, 0]
    out, s = _ss_run(ctx, c, A, B, 1.5, 0.5, seed)
    ref, nxt = O.sketch_general_sparse(left, layout, opS, opA, d, n, m, 1.5, c["S_rows"], c["S_cols"], 2, 1, 2, A, c["lda"], 0.5, B, c["ldb"],
                                       O.RNGState.from_words(seed))
    assert list(s.words()) == list(nxt.words())
    assert np.abs(out - ref).max() <= 1e-11 * np.abs(ref).max()


# ---- Axis::Long sparse operators (LASO, sparse_skops.hh:167-282, 669-704) ---------------------------------------------------------------------
from _qrcases import ls_input  # noqa: E402


@pytest.mark.parametrize("k", range(int(GS["la_count"])))
def test_fill_sparse_laso_golden(ctx, k):
    """COO export of an Axis::Long operator: indices, merged sqrt(count) * sign values (in the working precision), their order and the returned
    state are integer / correctly-rounded work: bit-exact against the golden triplets of the compiled reference."""
    r, c, nnz, sr, sc, ro, co = [int(x) for x in GS[f"la{k}_args"]]
    dt = torch.float64 if str(GS[f"la{k}_dtype"]) == "f64" else torch.float32
    D = rl.SparseDist(r, c, nnz, rl.AXIS_LONG)
    nz, vals, rows, cols, st = rl.fill_sparse(ctx, D, _st(GS[f"la{k}_state_in"]), dt, sub=(sr, sc, ro, co))
    assert nz == len(GS[f"la{k}_vals"]) and list(st.words()) == list(GS[f"la{k}_state_out"])
    assert np.array_equal(vals.cpu().numpy(), GS[f"la{k}_vals"])
    assert np.array_equal(rows.cpu().numpy(), GS[f"la{k}_rows"]) and np.array_equal(cols.cpu().numpy(), GS[f"la{k}_cols"])


@pytest.mark.parametrize("k", range(int(GS["ls_count"])))
def test_sketch_sparse_left_laso_golden(ctx, k):
    """Left sketch with a wide Axis::Long operator against golden outputs of the compiled reference (sub-matrix offsets, padded leading
    dimensions, alpha = 0.75, beta = -0.5): state exact, padding untouched, values 1e-12 (fp64) / 2e-5 (fp32) relative to the largest entry."""
    c = ls_input(k)
    dt = torch.float64 if c["dtype"] == np.float64 else torch.float32
    Ad = torch.from_numpy(c["A"]).cuda().view(c["n"], c["lda"]).t()[: c["m"], :]
    Bfull = torch.from_numpy(c["B"].copy()).cuda()
    Bd = Bfull.view(c["n"], c["ldb"]).t()[: c["d"], :]
    assert rl._ld(Ad) == c["lda"] and rl._ld(Bd) == c["ldb"] and Ad.dtype == dt
    s = _st(c["seed"])
    rl.sketch_general_left(ctx, rl.SparseDist(c["S_rows"], c["S_cols"], c["vec_nnz"], rl.AXIS_LONG), s, Ad, d=c["d"], alpha=0.75, beta=-0.5, B=Bd,
                           ro_s=c["ro"], co_s=c["co"])
    out, ref = Bfull.cpu().numpy(), GS[f"ls{k}_Bout"]
    assert list(s.words()) == list(GS[f"ls{k}_state_out"])
    mask = np.ones(ref.shape, dtype=bool)
    O._mat_view(mask, c["d"], c["n"], c["ldb"], O.LAYOUT_COLMAJOR)[:, :] = False
    assert np.array_equal(out[mask], c["B"][mask]), "padding of B was written"
    tol = 1e-12 if c["dtype"] == np.float64 else 2e-5
    assert np.abs(out - ref).max() <= tol * np.abs(ref).max()


def test_sketch_sparse_left_laso_vs_device_coo(ctx):
    """A larger case (d = 512, 2^16 x 96 data, vec_nnz = 16): the applied sketch equals the product with the operator exported by fill_sparse
    on the same state (1e-13 relative), and the state advances by min(S_rows, S_cols) * vec_nnz counters."""
    d, m, n, nnz = 512, 1 << 16, 96, 16
    D = rl.SparseDist(d, m, nnz, rl.AXIS_LONG)
    g = torch.Generator(device="cuda").manual_seed(3)
    A = rl.to_f(torch.randn((m, n), dtype=torch.float64, device="cuda", generator=g))
    s = rl.RNGState(5)
    B = rl.sketch_general_left(ctx, D, s, A)
    nz, vals, rows, cols, _ = rl.fill_sparse(ctx, D, rl.RNGState(5))
    S = torch.sparse_coo_tensor(torch.stack([rows, cols]), vals, (d, m)).coalesce()
    Bref = torch.sparse.mm(S, A.contiguous())
    assert (B - Bref).abs().max().item() <= 1e-13 * Bref.abs().max().item()
    assert nz <= d * nnz and list(s.words())[:1] == [d * nnz]
