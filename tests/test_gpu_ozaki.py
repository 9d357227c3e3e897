"""GPU: fp64 tall GEMMs through tcgen05 int8 digit slices (randlapack_b200/csrc/ozaki.cu) against torch fp64.

Tolerance (stated): 7 digits keep 48 bits below each row's (NN) / column-chunk's (TN) largest magnitude, so the error of one
output entry is bounded by ~K * 128^-7 * max|a_i.| * max|b_.j| ~ 2e-15 * K in those units; the test allows 1e-13 relative to
(|A| |B|)_ij, i.e. DGEMM-level componentwise-by-bound accuracy, including on inputs whose rows differ by 100 orders of magnitude."""
import numpy as np
import pytest
import torch

import randlapack_b200 as rl

pytestmark = pytest.mark.gpu


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
def test_i8_gemm_nn(ctx, shape, bad_scaling):
    m, K, N = shape
    A = _mk(m, K, 1, scale_rows=bad_scaling)
    B = _mk(K, N, 2, scale_cols=bad_scaling)
    C0 = _mk(m, N, 3)
    C = C0.clone()
    rl.gemm(ctx, False, False, -0.5, A, B, 2.0, C, engine="i8")
    ref = -0.5 * (A @ B) + 2.0 * C0
    bound = 0.5 * (A.abs() @ B.abs()) + 2.0 * C0.abs()
    err = ((C - ref).abs() / bound).max().item()
    assert err <= 1e-13, err


@pytest.mark.parametrize("shape", [(64, 128, 64), (1, 1, 1), (1000, 100, 17), (70000, 256, 64), (5000, 1024, 256), (40001, 130, 65)])
@pytest.mark.parametrize("bad_scaling", [False, True])
def test_i8_gemm_tn(ctx, shape, bad_scaling):
    m, N1, N2 = shape
    X = _mk(m, N1, 4, scale_cols=bad_scaling)
    Y = _mk(m, N2, 5, scale_cols=bad_scaling)
    C = rl.gemm(ctx, True, False, 1.0, X, Y, engine="i8")
    ref = X.t() @ Y
    bound = X.abs().t() @ Y.abs()
    err = ((C - ref).abs() / bound).max().item()
    assert err <= 1e-13, err
    # deterministic
    C2 = rl.gemm(ctx, True, False, 1.0, X, Y, engine="i8")
    assert torch.equal(C, C2)


def test_i8_gemm_matches_dmma_on_rsvd_shape(ctx):
    m, n, k = 1 << 18, 1024, 256
    A = _mk(m, n, 7)
    Om = _mk(n, k, 8)
    Y1 = rl.gemm(ctx, False, False, 1.0, A, Om, engine="i8")
    Y0 = rl.gemm(ctx, False, False, 1.0, A, Om, engine="dmma")
    assert ((Y1 - Y0).norm() / Y0.norm()).item() <= 1e-13
    Z1 = rl.gemm(ctx, True, False, 1.0, A, Y0, engine="i8")
    Z0 = rl.gemm(ctx, True, False, 1.0, A, Y0, engine="dmma")
    assert ((Z1 - Z0).norm() / Z0.norm()).item() <= 1e-13
