"""GPU: tall GEMMs through tcgen05 int8 digit slices (randlapack_b200/csrc/ozaki.cu) against torch fp64.

Tolerance (stated): S balanced base-256 digits keep 8S-2 bits below each row's (NN) / column-chunk's (TN) largest magnitude and
the digit pairs below 256^-(S-1) are dropped, so one output entry is off by at most ~K * S * 2^(-8S+4) * max|a_i.| * max|b_.j|.
Relative to (|A| |B|)_ij (componentwise-by-bound, including on inputs whose rows differ by 100 orders of magnitude) the tests
allow 1e-13 for S = 7 (DGEMM-level; observed ~1e-15), 2e-12 for S = 6 (the fast fp64 setting; observed ~1e-13) and 2e-7 for
fp32 storage with S = 4 (the fp32 output rounding itself is 6e-8)."""
import numpy as np
import pytest
import torch

import randlapack_b200 as rl

pytestmark = pytest.mark.gpu


TOL = {6: 2e-12, 7: 1e-13}


def _mk(m, n, seed, scale_rows=False, scale_cols=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    X = torch.randn((n, m), dtype=torch.float64, device="cuda", generator=g).t()
    if scale_rows:
        X = X * torch.pow(10.0, torch.linspace(-50, 50, m, dtype=torch.float64, device="cuda"))[:, None]
    if scale_cols:
        X = X * torch.pow(10.0, torch.linspace(-30, 30, n, dtype=torch.float64, device="cuda"))[None, :]
    return rl.to_f(X)


@pytest.mark.parametrize("shape", [(128, 32, 64), (1, 1, 1), (1000, 100, 17), (5000, 1024, 256), (40000, 300, 70), (333, 1031, 129)])
@pytest.mark.parametrize("bad_scaling", [False, True])
@pytest.mark.parametrize("digits", [6, 7])
def test_i8_gemm_nn(ctx, shape, bad_scaling, digits):
    m, K, N = shape
    ctx.set_i8_digits(digits)
    A = _mk(m, K, 1, scale_rows=bad_scaling)
    B = _mk(K, N, 2, scale_cols=bad_scaling)
    C0 = _mk(m, N, 3)
    C = C0.clone()
    rl.gemm(ctx, False, False, -0.5, A, B, 2.0, C, engine="i8")
    ref = -0.5 * (A @ B) + 2.0 * C0
    bound = 0.5 * (A.abs() @ B.abs()) + 2.0 * C0.abs()
    err = ((C - ref).abs() / bound).max().item()
    ctx.set_i8_digits(0)
    assert err <= TOL[digits], err


@pytest.mark.parametrize("shape", [(64, 128, 64), (1, 1, 1), (1000, 100, 17), (70000, 256, 64), (5000, 1024, 256), (40001, 130, 65)])
@pytest.mark.parametrize("bad_scaling", [False, True])
@pytest.mark.parametrize("digits", [6, 7])
def test_i8_gemm_tn(ctx, shape, bad_scaling, digits):
    m, N1, N2 = shape
    ctx.set_i8_digits(digits)
    X = _mk(m, N1, 4, scale_cols=bad_scaling)
    Y = _mk(m, N2, 5, scale_cols=bad_scaling)
    C = rl.gemm(ctx, True, False, 1.0, X, Y, engine="i8")
    ref = X.t() @ Y
    bound = X.abs().t() @ Y.abs()
    err = ((C - ref).abs() / bound).max().item()
    # deterministic
    C2 = rl.gemm(ctx, True, False, 1.0, X, Y, engine="i8")
    ctx.set_i8_digits(0)
    assert err <= TOL[digits], err
    assert torch.equal(C, C2)


@pytest.mark.parametrize("digits", [6, 7])
def test_i8_gemm_matches_dmma_on_rsvd_shape(ctx, digits):
    m, n, k = 1 << 18, 1024, 256
    A = _mk(m, n, 7)
    Om = _mk(n, k, 8)
    ctx.set_i8_digits(digits)
    Y1 = rl.gemm(ctx, False, False, 1.0, A, Om, engine="i8")
    Y0 = rl.gemm(ctx, False, False, 1.0, A, Om, engine="dmma")
    Z1 = rl.gemm(ctx, True, False, 1.0, A, Y0, engine="i8")
    Z0 = rl.gemm(ctx, True, False, 1.0, A, Y0, engine="dmma")
    ctx.set_i8_digits(0)
    assert ((Y1 - Y0).norm() / Y0.norm()).item() <= TOL[digits] / 10
    assert ((Z1 - Z0).norm() / Z0.norm()).item() <= TOL[digits] / 10


@pytest.mark.parametrize("shape", [(3000, 200, 70), (40000, 512, 128)])
def test_i8_gemm_f32_storage(ctx, shape):
    """fp32 storage: 4 digits (30 bits below the group maximum) against the fp64 product of the same fp32 inputs."""
    m, K, N = shape
    A = _mk(m, K, 11).float()
    A = rl.to_f(A)
    B = rl.to_f(_mk(K, N, 12).float())
    C = rl.gemm(ctx, False, False, 1.0, A, B, engine="i8")
    ref = A.double() @ B.double()
    bound = A.double().abs() @ B.double().abs()
    assert ((C.double() - ref).abs() / bound).max().item() <= 2e-7
    Y = rl.to_f(_mk(m, N, 13).float())
    Z = rl.gemm(ctx, True, False, 1.0, A, Y, engine="i8")
    refz = A.double().t() @ Y.double()
    boundz = A.double().abs().t() @ Y.double().abs()
    assert ((Z.double() - refz).abs() / boundz).max().item() <= 2e-7


@pytest.mark.parametrize("digits", [6, 7])
def test_i8_gemm_column_graded_first_operand_is_normwise(ctx, digits):
    """the per-row scale of the NN product cannot remove a grading ALONG a row.  With the columns of A spread over
    14 decades an entry 1e-14 below its row maximum keeps none of its bits: the scheme is fp64-accurate NORM-wise per row
    (|C - AB|_ij <= K 2^-(8S-3) max_k|a_ik| max_k|b_kj|), not component-wise in |A||B|.  This test pins exactly that statement: the
    normwise bound holds, and the componentwise-by-|A||B| error of a product that is dominated by the small columns is NOT small."""
    m, K, N = 20000, 256, 128
    A = _mk(m, K, 21, scale_cols=False)
    A = rl.to_f(A * torch.pow(10.0, torch.linspace(0, -14, K, dtype=torch.float64, device="cuda"))[None, :])
    B = _mk(K, N, 22)
    ctx.set_i8_digits(digits)
    C = rl.gemm(ctx, False, False, 1.0, A, B, engine="i8")
    ctx.set_i8_digits(0)
    ref = A @ B
    bound = A.abs().max(dim=1).values[:, None] * B.abs().max(dim=0).values[None, :] * K
    P = 8 * digits - 2
    # 2^-(P-1): both operands rounded to P bits below their group maximum; 2^-52: the fp64 rounding of the result and of the torch reference
    tol = 2.0 ** -(P - 1) + 2.0 ** -52
    assert ((C - ref).abs() / bound).max().item() <= tol
    # a right-hand side that only sees the tiny columns: the product is ~1e-14 of the row scale and inherits the absolute error
    B2 = B.clone()
    B2[: K - 8] = 0
    C2 = rl.gemm(ctx, False, False, 1.0, A, B2, engine="i8")
    ref2 = A @ B2
    bound2 = A.abs().max(dim=1).values[:, None] * B2.abs().max(dim=0).values[None, :] * K
    assert ((C2 - ref2).abs() / bound2).max().item() <= tol


@pytest.mark.parametrize("bad", [float("nan"), float("inf")])
def test_i8_gemm_nonfinite_propagates(ctx, bad):
    """a NaN / Inf entry of an operand must reach the output (the reference BLAS path propagates it), not turn into finite
    garbage.  Rows (NN) / columns (TN) that contain the entry come back non-finite; the others stay accurate."""
    m, K, N = 20000, 256, 128
    A = _mk(m, K, 31)
    B = _mk(K, N, 32)
    A[777, 5] = bad
    C = rl.gemm(ctx, False, False, 1.0, A, B, engine="i8")
    assert not torch.isfinite(C[777]).any()
    ok = torch.ones(m, dtype=torch.bool, device="cuda"); ok[777] = False
    Az = A.clone(); Az[777] = 0
    assert ((C[ok] - (Az @ B)[ok]).abs().max() / (Az.abs() @ B.abs()).max()).item() <= 2e-12
    Y = _mk(m, N, 33)
    Z = rl.gemm(ctx, True, False, 1.0, A, Y, engine="i8")
    assert not torch.isfinite(Z[5]).any()
    okc = torch.ones(K, dtype=torch.bool, device="cuda"); okc[5] = False
    assert torch.isfinite(Z[okc]).all()


# ---------------------------------------------------------------------------------------------------------------------------
# fused engine (ozaki_fused.cu): the digits of the tall operand are produced inside the tensor-core kernel
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(20000, 1024, 256), (70001, 300, 130), (64, 32, 96), (1, 1, 100), (4000, 1031, 129), (333, 77, 384)])
@pytest.mark.parametrize("bad_scaling", [False, True])
@pytest.mark.parametrize("digits", [6, 7])
def test_i8_fused_gemm_nn(ctx, shape, bad_scaling, digits):
    """A(m x K) B(K x N), N >= 96 (the fused engine's range): ragged m (not a multiple of 64), K (not of 32) and N (not of 128), alpha/beta,
    against torch fp64 with the componentwise-by-bound tolerance of the staged engine, and against the staged engine itself."""
    m, K, N = shape
    ctx.set_i8_digits(digits)
    A = _mk(m, K, 41, scale_rows=bad_scaling)
    B = _mk(K, N, 42, scale_cols=bad_scaling)
    C0 = _mk(m, N, 43)
    try:
        ctx.set_i8_fused(True)
        C = C0.clone()
        rl.gemm(ctx, False, False, -0.5, A, B, 2.0, C, engine="i8")
        ctx.set_i8_fused(False)
        Cs = C0.clone()
        rl.gemm(ctx, False, False, -0.5, A, B, 2.0, Cs, engine="i8")
    finally:
        ctx.set_i8_fused(True)
        ctx.set_i8_digits(0)
    ref = -0.5 * (A @ B) + 2.0 * C0
    bound = 0.5 * (A.abs() @ B.abs()) + 2.0 * C0.abs()
    assert ((C - ref).abs() / bound).max().item() <= TOL[digits]
    assert ((C - Cs).abs() / bound).max().item() <= 2 * TOL[digits]


@pytest.mark.parametrize("shape", [(70000, 256, 128), (40001, 130, 97), (5000, 1024, 256), (100, 64, 96), (1200000, 64, 128)])
@pytest.mark.parametrize("bad_scaling", [False, True])
@pytest.mark.parametrize("digits", [6, 7])
def test_i8_fused_gemm_tn(ctx, shape, bad_scaling, digits):
    """X(m x N1)^T Y(m x N2), N2 >= 96: several 16384-row accumulation groups, a ragged last group, more groups than one launch takes
    (1.2 M rows: 74 groups, 64 per launch), deterministic."""
    m, N1, N2 = shape
    ctx.set_i8_digits(digits)
    X = _mk(m, N1, 44, scale_cols=bad_scaling)
    Y = _mk(m, N2, 45, scale_cols=bad_scaling)
    try:
        C = rl.gemm(ctx, True, False, 1.0, X, Y, engine="i8")
        C2 = rl.gemm(ctx, True, False, 1.0, X, Y, engine="i8")
    finally:
        ctx.set_i8_digits(0)
    ref = X.t() @ Y
    bound = X.abs().t() @ Y.abs()
    assert ((C - ref).abs() / bound).max().item() <= TOL[digits]
    assert torch.equal(C, C2)


def test_i8_fused_gemm_f32_storage(ctx):
    m, K, N = 40000, 512, 128
    A = rl.to_f(_mk(m, K, 51).float())
    B = rl.to_f(_mk(K, N, 52).float())
    C = rl.gemm(ctx, False, False, 1.0, A, B, engine="i8")
    ref = A.double() @ B.double()
    bound = A.double().abs() @ B.double().abs()
    assert ((C.double() - ref).abs() / bound).max().item() <= 2e-7
    Y = rl.to_f(_mk(m, N, 53).float())
    Z = rl.gemm(ctx, True, False, 1.0, A, Y, engine="i8")
    refz = A.double().t() @ Y.double()
    boundz = A.double().abs().t() @ Y.double().abs()
    assert ((Z.double() - refz).abs() / boundz).max().item() <= 2e-7


@pytest.mark.parametrize("shape", [(70001, 256), (20000, 128), (33000, 384), (16500, 512)])
@pytest.mark.parametrize("digits", [6, 7])
def test_i8_fused_gemm_nn_inplace(ctx, shape, digits):
    """U = Y M in place (rl_rsvd.hh:148 with U stored over Y): the ceil(N / 128) CTAs of a row tile form a cluster and synchronise between
    their last read and their first write (1, 2 or 4 column tiles; 384 columns = 3 tiles take the staged engine).  Against torch fp64 and
    against the out-of-place product (bit-identical)."""
    m, k = shape
    ctx.set_i8_digits(digits)
    Y = _mk(m, k, 61)
    M = _mk(k, k, 62)
    try:
        C = rl.gemm(ctx, False, False, 1.0, Y, M, engine="i8")
        U = Y.clone()
        rl.gemm(ctx, False, False, 1.0, U, M, 0.0, U, engine="i8")
    finally:
        ctx.set_i8_digits(0)
    ref = Y @ M
    bound = Y.abs() @ M.abs()
    assert ((U - ref).abs() / bound).max().item() <= TOL[digits]
    assert torch.equal(U, C)
