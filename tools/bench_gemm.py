#!/usr/bin/env python
"""Time the two tall fp64 products of the RSVD (Y = A Omega, Z = A^T Y) on both engines: DMMA and tcgen05 int8 digit slices.
usage: python tools/bench_gemm.py [log2 m] [n] [k]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import randlapack_b200 as rl  # noqa: E402

lm = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
k = int(sys.argv[3]) if len(sys.argv) > 3 else 256
only = sys.argv[4] if len(sys.argv) > 4 else ""   # "nn" / "tn": only that product on the i8 engine (6 digits); "i8": both, i8 only
m = 1 << lm
ctx = rl.Context(0)
if os.environ.get("RLB200_I8_FUSED") == "0":
    ctx.set_i8_fused(False)
dev = torch.device("cuda", 0)
A = rl.empty_f(m, n, torch.float64, dev)
ctx.check(ctx._lib.rlb200_fill_dense_f64_dev(ctx._h, m, n, 0, 0, 0, m, n, 0, 0, A.data_ptr(), rl.RNGState(1).words()))
Om, _ = rl.fill_dense(ctx, rl.DenseDist(n, k), rl.RNGState(2))
Om = Om.view(k, n).t()
Y = rl.empty_f(m, k, torch.float64, dev)
Z = rl.empty_f(n, k, torch.float64, dev)
out = {"m": m, "n": n, "k": k}
names = ["gemm_nn", "gemm_tn", "rightmul", "small", "fill", "sketch", "factor"]
for engname in (("i8s6",) if only else ("dmma", "i8s6", "i8s7")):
    eng = "dmma" if engname == "dmma" else "i8"
    ctx.set_i8_digits(int(engname[-1]) if eng == "i8" else 0)
    for op in (("nn", "tn") if only in ("", "i8") else (only,)):
        def run():
            if op == "nn":
                rl.gemm(ctx, False, False, 1.0, A, Om, 0.0, Y, engine=eng)
            else:
                rl.gemm(ctx, True, False, 1.0, A, Y, 0.0, Z, engine=eng)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        ctx.timers_enable(True)
        for w in range(7):
            ctx.timer_read(w, reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tm = {nm: round(ctx.timer_read(i)[0] / reps, 3) for i, nm in enumerate(names)}
        ctx.timers_enable(False)
        out[f"{engname}_{op}"] = {"ms": round(ms, 3), "tflops": round(2.0 * m * n * k / ms / 1e9, 2), "timers_ms": {a: b for a, b in tm.items() if b}}
print(json.dumps(out))
