#!/usr/bin/env python
"""Debug: RSVD on the headline engine path (k >= 64), fused engine on/off: sigma error, orthogonality of U and V vs the oracle."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import randlapack_b200 as rl
from oracle import rl_oracle as O
import _ref

def dev(a): return torch.from_numpy(np.ascontiguousarray(a.T)).cuda().t()
def host(t): return np.asfortranarray(t.detach().cpu().numpy())
ctx = rl.Context(0)
m, n = 32768, 512
A, st0 = O.gen_poly_mat(m, n, n, 2025.0, 2.0, O.RNGState(0))
Ad = dev(A)
for k, p in [(256, 2), (256, 3), (128, 2), (256, 0)]:
    st_d = rl.RNGState(st0.key, st0.counter)
    rows = n if p % 2 == 0 else m
    buf, _ = rl.fill_dense(ctx, rl.DenseDist(rows, k), st_d.copy())
    Om = np.asfortranarray(buf.cpu().numpy().reshape((rows, k), order="F"))
    *_, rsvd_o = O.make_stack(O.StackOpts(p, 1, k, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ))
    rc_o, kk_o, U_o, S_o, V_o, s_o = rsvd_o.call(A, k, 0.0, st0.copy(), omega_override=Om)
    for fused in (True, False):
        for gf in (True, False):
            if gf: os.environ.pop("RLB200_NO_GRAM_FUSION", None)
            else: os.environ["RLB200_NO_GRAM_FUSION"] = "1"
            ctx.set_i8_fused(fused)
            stack = rl.RSVD(rl.QB(rl.RF(rl.RS(rl.CholQRQ(), p, 1), rl.CholQRQ()), rl.CholQRQ()), k)
            rc, kk, U, S, V = stack.call(ctx, Ad, k, 0.0, st_d.copy())
            U, S, V = host(U)[:, :kk], S.cpu().numpy()[:kk], host(V)[:, :kk]
            print(f"k={k} p={p} fused={fused} gramfusion={gf}: rc={rc} kk={kk} sig_err={np.abs(S-S_o).max()/S_o[0]:.2e} "
                  f"orthU={np.linalg.norm(U.T@U-np.eye(kk)):.2e} orthV={np.linalg.norm(V.T@V-np.eye(kk)):.2e} "
                  f"orthU_ref={np.linalg.norm(U_o.T@U_o-np.eye(kk)):.2e} sin={_ref.subspace_sin(U_o, U):.2e}", flush=True)
os.environ.pop("RLB200_NO_GRAM_FUSION", None)
# the Gram output of the fused TN product itself, against torch
torch.manual_seed(1)
for mm, n1, n2 in [(70000, 512, 256), (40000, 300, 128)]:
    X = torch.randn(n1, mm, dtype=torch.float64, device="cuda").t()
    Y = (torch.randn(n2, mm, dtype=torch.float64, device="cuda") * torch.logspace(0, -3, n2, dtype=torch.float64, device="cuda")[:, None]).t()
    if hasattr(rl, "gemm_tn_gram"):
        Z, G = rl.gemm_tn_gram(ctx, X, Y)
        refG = Y.t() @ Y; refZ = X.t() @ Y
        bG = Y.abs().t() @ Y.abs(); bZ = X.abs().t() @ Y.abs()
        eu = torch.triu((G - refG).abs() / bG).max().item()
        print(f"tn_gram {mm}x{n1}x{n2}: Z err {((Z-refZ).abs()/bZ).max().item():.2e}  G(upper) err {eu:.2e}", flush=True)
