#!/usr/bin/env python
"""Debug: fused NN product with digits shared across a cluster; prints the error and where the wrong entries are."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import randlapack_b200 as rl
m, K, N, S = [int(x) for x in sys.argv[1:5]]
ctx = rl.Context(0)
g = torch.Generator(device="cuda").manual_seed(1)
A = rl.to_f(torch.randn((K, m), dtype=torch.float64, device="cuda", generator=g).t())
B = rl.to_f(torch.randn((N, K), dtype=torch.float64, device="cuda", generator=g).t())
ctx.set_i8_digits(S)
C = rl.gemm(ctx, False, False, 1.0, A, B, engine="i8")
torch.cuda.synchronize()
ref = A @ B
bound = A.abs() @ B.abs()
err = (C - ref).abs() / bound
bad = err > 1e-10
print(f"m={m} K={K} N={N} S={S}: max err {err.max().item():.3e}, wrong entries {int(bad.sum())} of {m*N}")
if bad.any():
    rows = bad.any(dim=1).nonzero().flatten(); cols = bad.any(dim=0).nonzero().flatten()
    print(" wrong rows:", rows[:20].tolist(), "... count", len(rows), " tiles(64):", sorted(set((rows // 64).tolist()))[:20])
    print(" wrong cols:", cols[:20].tolist(), "... count", len(cols), " tiles(128):", sorted(set((cols // 128).tolist())))
