"""Time geqp3 / geqrf of a d x n device matrix (rl.qr_small): python tools/bench_qrcp.py d n [f32|f64] [pivot 0/1]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import randlapack_b200 as rl
d, n = int(sys.argv[1]), int(sys.argv[2])
dt = torch.float32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else torch.float64
pivot = (len(sys.argv) <= 4) or sys.argv[4] == "1"
ctx = rl.Context(0)
g = torch.Generator(device="cuda").manual_seed(1)
A0 = rl.to_f(torch.randn((d, n), dtype=dt, device="cuda", generator=g) * (1.0 + torch.arange(n, device="cuda", dtype=dt))[None, :] ** -0.5)
ts = []
for it in range(4):
    A = A0.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    J, tau = rl.qr_small(ctx, A, pivot)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
# check: R^T R == (A P)^T (A P)
R = torch.triu(A[:min(d, n), :])
AP = A0[:, (J - 1)] if pivot else A0
err = ((R.t() @ R) - (AP.t() @ AP)).abs().max().item() / (AP.t() @ AP).abs().max().item()
print({"d": d, "n": n, "dtype": str(dt), "pivot": pivot, "ms": [round(t, 3) for t in ts], "gram_err": err,
       "coop": os.environ.get("RLB200_QR_NOCOOP") is None})
