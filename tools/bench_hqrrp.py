"""Time hqrrp (rl.hqrrp) and BQRRP on the same m x n device matrix, and CQRRPT with qrcp = geqp3 / hqrrp:
python tools/bench_hqrrp.py m n nb pp [f32|f64].  Prints one JSON line per measurement (wall clock around a synchronised call: the driver
returns after its final device-to-host copy of the pivots)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import randlapack_b200 as rl

m, n, nb, pp = (int(x) for x in sys.argv[1:5])
dt = torch.float32 if (len(sys.argv) > 5 and sys.argv[5] == "f32") else torch.float64
ctx = rl.Context(0)
g = torch.Generator(device="cuda").manual_seed(1)
A0 = rl.to_f(torch.randn((m, n), dtype=dt, device="cuda", generator=g) * (1.0 + torch.arange(n, device="cuda", dtype=dt))[None, :] ** -0.5)
flop = 2.0 * m * n * n - 2.0 * n ** 3 / 3.0 if m >= n else 2.0 * n * m * m - 2.0 * m ** 3 / 3.0


def timed(fn, reps=3):
    ts = []
    for _ in range(reps):
        A = A0.clone()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn(A)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return ts, A, out


def check(A, tau, J):
    k = min(m, n, 2048)
    R = torch.triu(A[:k, :k]).double()
    AP = A0[:, (J[:k] - 1)].double()
    G = AP.t() @ AP
    return ((R.t() @ R) - G).abs().max().item() / G.abs().max().item()


for piv, qt, name in ((1, 0, "hqrrp_panel_pivoting"), (0, 1, "hqrrp_geqrf_panels"), (0, 2, "hqrrp_cholqr_panels")):
    ts, A, (rc, tau, J) = timed(lambda A: rl.hqrrp(ctx, A, nb, pp, piv, qt, rl.RNGState(0)))
    print(json.dumps({"what": name, "m": m, "n": n, "nb_alg": nb, "pp": pp, "dtype": str(dt), "ms": [round(t, 2) for t in ts],
                      "tflops": round(flop / min(ts) / 1e9, 2), "rc": rc, "gram_err_leading_block": check(A, tau, J)}), flush=True)
for qt, name in ((rl.QRTALL_CHOLQR, "bqrrp_cholqr"), (rl.QRTALL_GEQRF, "bqrrp_geqrf")):
    alg = rl.BQRRP(False, nb)
    alg.qr_tall = qt
    ts, A, (rc, tau, J) = timed(lambda A: alg.call(ctx, A, 1.0 + pp / nb, rl.RNGState(0)))
    print(json.dumps({"what": name, "m": m, "n": n, "b": nb, "dtype": str(dt), "ms": [round(t, 2) for t in ts],
                      "tflops": round(flop / min(ts) / 1e9, 2), "rc": rc, "rank": alg.rank, "gram_err_leading_block": check(A, tau, J)}), flush=True)
