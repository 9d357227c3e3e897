"""GPU parity of the tall-skinny GEMM kernels (DMMA) against a plain torch fp64 reference of the same op.
Tolerance: fp64 products of O(1) entries with contraction length K: |err| <= 64 * K * eps * max|A| max|B|
(fp32 storage: same bound with eps_f32 for the final rounding)."""
import pytest
import torch

import randlapack_b200 as rl

pytestmark = pytest.mark.gpu


def _f(m, n, dtype, gen, ld=None):
    ld = ld or m
    buf = torch.randn((n, ld), dtype=torch.float64, device="cuda", generator=gen).to(dtype)
    return buf.t()[:m, :] if ld != m else buf.t()


SHAPES_NN = [(1000, 256, 64), (4096, 32, 256), (1, 1, 1), (129, 17, 5), (777, 130, 33), (5000, 256, 1024), (64, 200, 7), (300, 64, 100)]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("m,N,K", SHAPES_NN)
def test_gemm_nn(ctx, dtype, m, N, K):
    g = torch.Generator(device="cuda").manual_seed(m + N + K)
    for pad in (0, 3):
        A, B = _f(m, K, dtype, g, m + pad), _f(K, N, dtype, g, K + pad)
        C0 = _f(m, N, dtype, g)
        C = C0.clone().t().contiguous().t()
        rl.gemm(ctx, False, False, 1.5, A, B, -0.5, C)
        ref = 1.5 * (A.double() @ B.double()) - 0.5 * C0.double()
        eps = 2.2e-16 if dtype == torch.float64 else 1.2e-7
        tol = 64 * K * 2.2e-16 * 25 + (0 if dtype == torch.float64 else eps * ref.abs().max().item() * 2)
        assert (C.double() - ref).abs().max().item() <= tol


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("m,N1,N2", [(100000, 256, 64), (5000, 1024, 256), (37, 5, 3), (20000, 130, 70), (1, 4, 4), (4097, 64, 64), (70000, 32, 32)])
def test_gemm_tn(ctx, dtype, m, N1, N2):
    g = torch.Generator(device="cuda").manual_seed(m + N1)
    for pad in (0, 1):
        A, B = _f(m, N1, dtype, g, m + pad), _f(m, N2, dtype, g, m + pad)
        C = rl.gemm(ctx, True, False, 1.0, A, B)
        ref = A.double().t() @ B.double()
        eps = 2.2e-16 if dtype == torch.float64 else 1.2e-7
        tol = 64 * (m ** 0.5 + 16) * 2.2e-16 * 25 + (0 if dtype == torch.float64 else eps * ref.abs().max().item() * 2)
        assert (C.double() - ref).abs().max().item() <= tol
        C2 = rl.gemm(ctx, True, False, 1.0, A, B)
        assert torch.equal(C, C2), "split-K reduction must be run-to-run deterministic"


@pytest.mark.parametrize("m,N,K", [(3000, 200, 16), (515, 1024, 8), (100, 7, 3)])
def test_gemm_nt(ctx, m, N, K):
    g = torch.Generator(device="cuda").manual_seed(5)
    A, B = _f(m, K, torch.float64, g), _f(N, K, torch.float64, g)
    C0 = _f(m, N, torch.float64, g)
    C = C0.clone().t().contiguous().t()
    rl.gemm(ctx, False, True, -1.0, A, B, 1.0, C)
    ref = C0 - A @ B.t()
    assert (C - ref).abs().max().item() <= 64 * K * 2.2e-16 * 25


def test_gemm_linearity_large(ctx):
    # size-independent property at a large size: A(x + y) == Ax + Ay to round-off, 2^21 x 512 by 512 x 64
    g = torch.Generator(device="cuda").manual_seed(9)
    m, K, N = 1 << 21, 512, 64
    A = _f(m, K, torch.float64, g)
    X, Y = _f(K, N, torch.float64, g), _f(K, N, torch.float64, g)
    lhs = rl.gemm(ctx, False, False, 1.0, A, (X + Y).t().contiguous().t())
    rhs = rl.gemm(ctx, False, False, 1.0, A, X) + rl.gemm(ctx, False, False, 1.0, A, Y)
    assert (lhs - rhs).abs().max().item() <= 1e-10
    # and against torch on a row sample
    idx = torch.randint(0, m, (512,), device="cuda", generator=g)
    assert (lhs[idx] - A[idx] @ (X + Y)).abs().max().item() <= 1e-10
