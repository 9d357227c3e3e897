#!/usr/bin/env python
"""bench.py — rank-k RSVD throughput (Gflop/s) on tall dense fp64 A, the metric BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--m M --n N --k K --p P]

One "step" = one complete RSVD (RS -> RF -> CholQRQ -> QB -> SVD(B) -> U) of a resident synthetic A.
N = 1: BASELINE.json configs[1] (2^24 x 1024 fp64, rank 256).  N > 1 (torchrun): the same per-GPU row block on every
rank (weak scaling), i.e. an (N * 2^24) x 1024 row-sharded A with Gram / B^T / norm allreduces over NCCL.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of ALL kernels of one step per row of A at n = 1024, k = 256, p = 2, from the
# ncu launch list of this very command at m = 2^20 (profiles/launches_dram_r2.csv: every launch of one step, summed); and of the two
# dominant launches alone from their `ncu --set full` captures (profiles/ncu_oz2_nn_r2.txt: 8.598 GB + 2.092 GB per 2^20 rows;
# profiles/ncu_oz2_tn_r2.txt: 4.485 GB + 0.078 GB per 27 x 16384 rows).  Algorithmic bytes per row: 8 n = 8192 (A read once per pass).
OZ2_TRAFFIC_NN_PER_ROW = (8.597615e9 + 2.091834e9) / float(1 << 20)
OZ2_TRAFFIC_TN_PER_ROW = (4.485259e9 + 0.077737e9) / (27.0 * 16384.0)
# whole step, ALL kernels (profiles/launches_dram_r2b.csv: the ncu launch list of `bench.py --steps 1 --warmup 1 --m 1048576`, 3 steps executed:
# 242.18 GB read + written by this library's kernels = 80.73 GB per step of 2^20 rows; the persistent A*Omega kernel reads 1.47x the bytes of A
# from DRAM - L2 misses on re-referenced lines)
STEP_TRAFFIC_PER_ROW = 80.72608972799999e9 / float(1 << 20)
OZ_TRAFFIC_SOURCE = ("traffic: every kernel of one step, ncu launch list of this command at m = 2^20 (profiles/launches_dram_r2b.csv) scaled by m; "
                     "traffic_dominant_launches_only: ncu --set full captures of oz2_kernel (profiles/ncu_oz2_nn_r2.txt, ncu_oz2_tn_r2.txt), DRAM bytes per row "
                     "of A x rows x launches; algorithmic bytes per row and pass: 8 n (A, read once as fp64) + 8 k (Y written) for A*Omega, "
                     "8 n + S k (digits of Y) for A^T*Y")
# The contract is ONE JSON line on stdout.  Libraries loaded below may print there too (NCCL's version banner when a communicator is
# created): the process's stdout is pointed at stderr for the whole run and the JSON line is written to the saved descriptor.
_REAL_STDOUT = None


def _capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, line)
    else:
        os.write(_REAL_STDOUT, line)


METRIC = "rsvd_gflops"
UNIT = "Gflop/s"


def rsvd_flops(m, n, k, p, q=1):
    """Algorithmic flop count (SURVEY.md 8d / DESIGN.md): (p+2) tall GEMMs of 2mnk, CholQR of Q (2mk^2), U = Q*W (2mk^2),
    plus one CholQR per stabilised power pass (2mk^2 on the m x k iterate, 2nk^2 on the n x k one).  The O(nk^2 + k^3) small
    factorizations are not counted."""
    f = (p + 2) * 2.0 * m * n * k + 4.0 * m * k * k
    for pass_idx in range(1, p + 1):
        if pass_idx % q == 0:
            # with p even the passes alternate A*Omega (m x k) and A^T*Omega_1 (n x k), starting with m x k
            # with p odd the first pass is A^T*Omega_1 (n x k)
            tall = (pass_idx % 2 == 1) if p % 2 == 0 else (pass_idx % 2 == 0)
            f += 2.0 * (m if tall else n) * k * k
    return f


def class_flops(m, n, k, p, q=1, folded=True):
    """Flops EXECUTED per kernel class for one step: NN (A*Omega), TN (A^T*Y and the Gram syrk's), RIGHTMUL (Y*R^-1, Q*W).
    With folded solves (the default: A^T(Y R^-1) computed as (A^T Y) R^-1, Q W as Y (R^-1 W)) the m x k triangular solves of
    CholQR are not executed; the n x k / k x k ones that replace them are O(n k^2) and not counted."""
    n_even = p // 2
    nn = (n_even + 1) * 2.0 * m * n * k
    tn = (p - n_even + 1) * 2.0 * m * n * k
    rm = 2.0 * m * k * k        # U = Q W
    tn += 1.0 * m * k * k       # syrk of CholQR(Q)
    if not folded:
        rm += 1.0 * m * k * k   # trsm of CholQR(Q)
    for pass_idx in range(1, p + 1):
        if pass_idx % q == 0:
            tall = (pass_idx % 2 == 1) if p % 2 == 0 else (pass_idx % 2 == 0)
            d = m if tall else n
            tn += 1.0 * d * k * k
            if not (folded and tall):
                rm += 1.0 * d * k * k
    return {"gemm_nn": nn, "gemm_tn": tn, "rightmul": rm}


class ClockSampler:
    """nvidia-smi sampling of SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_fp64_peak():
    """fp64 tensor-pipe (DMMA.8x8x4) peak of THIS device from tools/peaks (register-resident loop, ~1 s).
    MEASURED_PEAKS.json only has HBM and bf16, so the fp64 denominator is measured here and labelled as such."""
    exe = os.path.join(ROOT, "tools", "peaks")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout.strip().splitlines()[-1]
        d = json.loads(out)
        return max(d["dmma_tflops"], d["dfma_tflops"]), "measured in-run by tools/peaks (DMMA.8x8x4 / DFMA register loop)", d
    except Exception as e:  # noqa: BLE001
        return 37.0, f"nominal B200 fp64 (tools/peaks failed: {type(e).__name__})", None


def measured_i8_peak():
    """int8 tensor-pipe peak of THIS device from tools/peaks_i8 (tcgen05.mma.kind::i8, M = 128, N = 256, K = 32 on resident operands; burst
    and ~1.5 s sustained), measured in-run: MEASURED_PEAKS.json has no int8 figure.  Also returns the cost of the engine's own instruction
    mixes (8 stacked instructions per K step for 6 digits)."""
    exe = os.path.join(ROOT, "tools", "peaks_i8")
    try:
        out = subprocess.run([exe, "1.5", "quick"], capture_output=True, text=True, timeout=180).stdout.strip().splitlines()[-1]
        d = json.loads(out)
        r = {x["pattern"]: x for x in d["results"] if x["cta_group"] == 1}
        return r["n256"]["burst_tops"], r["n256"]["sustained_tops"], "measured in-run by tools/peaks_i8 (UTCIMMA M=128 N=256 K=32, operands resident)", r
    except Exception as e:  # noqa: BLE001
        return None, None, f"tools/peaks_i8 failed: {type(e).__name__}", None


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU RSVD (oracle/_ref when it was compiled, else the oracle port)
# --------------------------------------------------------------------------------------------------
def cpu_rsvd_runner(n, k, p, q):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _ref
    from oracle import rl_oracle as O
    o = O.StackOpts(p, q, k, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ, False, False)
    R = _ref.ref_lib()
    cores = os.cpu_count() or 1
    if R is not None:
        R.rlref_set_num_threads(cores)

        def run(A):
            rc, kk, U, S, V, st = _ref.ref_rsvd(R, A, k, 0.0, [0] * 6, o)
            return S
        return run, "reference", cores
    *_, rsvd = O.make_stack(o)

    def run(A):
        rc, kk, U, S, V, st = rsvd.call(A, k, 0.0, O.RNGState(0))
        return S
    return run, "port", cores


def cpu_sample(n, k, p, q, m_cpu, steps, warmup):
    import numpy as np
    from oracle import rl_oracle as O
    run, kind, cores = cpu_rsvd_runner(n, k, p, q)
    # the same synthetic input family as the GPU arm: a Philox/Box-Muller DenseDist(m, n) sample (RandBLAS fill_dense, oracle restatement)
    A, _ = O.fill_dense(m_cpu, n, O.RNGState(0xA2))
    A = np.asfortranarray(A)
    for _ in range(warmup):
        run(A)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        run(A)
        ts.append(time.perf_counter() - t0)
    t = sum(ts) / len(ts)
    gf = rsvd_flops(m_cpu, n, k, p, q) / t / 1e9
    return gf, t, kind, cores


def parity_sample(ctx, rl, n, k, p, q, m_par):
    """Parity measured in the same run: RSVD of an m_par x n Philox matrix on the device and by the CPU restatement of the
    reference (oracle/, validated against the compiled reference) on the SAME operator Omega (the device's; the reference's own CPU-vs-GPU
    test feeds one sketch to both sides, test_bqrrp_gpu.cu:91-103).  Checker only - nothing here is timed."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _ref
    from oracle import rl_oracle as O
    dev = torch.device("cuda", torch.cuda.current_device())
    A = rl.empty_f(m_par, n, torch.float64, dev)
    ctx.check(ctx._lib.rlb200_fill_dense_f64_dev(ctx._h, m_par, n, rl.FAMILY_GAUSSIAN, rl.AXIS_LONG, rl.LAYOUT_NATURAL, m_par, n, 0, 0,
                                                 A.data_ptr(), rl.RNGState(0xA2).words()))
    # a decaying spectrum (sigma_j = j^-1) so that the rank-k truncation, the residual and the subspace are well defined
    A *= (1.0 / torch.arange(1, n + 1, dtype=torch.float64, device=dev))[None, :]
    st = rl.RNGState(7)
    rows = n if p % 2 == 0 else m_par
    buf, _ = rl.fill_dense(ctx, rl.DenseDist(rows, k), st.copy())
    Om = np.asfortranarray(buf.cpu().numpy().reshape((rows, k), order="F"))
    stack = rl.RSVD(rl.QB(rl.RF(rl.RS(rl.CholQRQ(), p, q), rl.CholQRQ()), rl.CholQRQ()), k)
    rc, kk, U, S, V = stack.call(ctx, A, k, 0.0, st.copy())
    Ah = np.asfortranarray(A.cpu().numpy())
    *_, rsvd_o = O.make_stack(O.StackOpts(p, q, k, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ, O.STAB_CHOLQRQ))
    rc_o, kk_o, U_o, S_o, V_o, _ = rsvd_o.call(Ah, k, 0.0, O.RNGState(7), omega_override=Om)
    U, S, V = U.cpu().numpy()[:, :kk], S.cpu().numpy()[:kk], V.cpu().numpy()[:, :kk]
    nrm = np.linalg.norm(Ah)
    r_dev = np.linalg.norm(Ah - (U * S) @ V.T) / nrm
    r_ref = np.linalg.norm(Ah - (U_o * S_o) @ V_o.T) / nrm
    return {"sample": f"{m_par} x {n} fp64, k={k}, p={p}: device RSVD vs the CPU restatement of the reference on the same operator",
            "codes_equal": bool((rc, kk) == (rc_o, kk_o)), "sigma_max_rel_err": float(np.abs(S - S_o).max() / S_o[0]),
            "residual_rel": {"device": float(r_dev), "reference": float(r_ref), "abs_diff": float(abs(r_dev - r_ref))},
            "subspace_sin": float(_ref.subspace_sin(U_o, U)), "orth_U": float(np.linalg.norm(U.T @ U - np.eye(kk))),
            "tolerance": "sigma 1e-10 relative, residual within 1e-10, subspace sin 1e-9 (north star / tests/test_gpu_drivers.py)"}


# --------------------------------------------------------------------------------------------------
# secondary workloads (BASELINE.json configs[2], configs[3] and the sketch micro-kernels): not the headline line; run by hand,
# results kept under profiles/.  Same timing rules: CUDA events, warm-up >= 3, inputs far larger than L2.
# --------------------------------------------------------------------------------------------------
def run_secondary(args):
    import torch
    import randlapack_b200 as rl
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    ctx = rl.Context(0)
    ctx.set_fp64_engine(args.engine)
    ctx.set_i8_digits(args.digits)
    peaks = load_peaks()
    hbm = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    wl = args.workload
    fill64, fill32 = ctx._lib.rlb200_fill_dense_f64_dev, ctx._lib.rlb200_fill_dense_f32_dev

    def gen(m, n, dtype, key):
        A = rl.empty_f(m, n, dtype, dev)
        fn = fill64 if dtype == torch.float64 else fill32
        ctx.check(fn(ctx._h, m, n, rl.FAMILY_GAUSSIAN, rl.AXIS_LONG, rl.LAYOUT_NATURAL, m, n, 0, 0, A.data_ptr(), rl.RNGState(key).words()))
        return A

    def timed(step, prep, steps, warmup):
        ts = []
        for i in range(warmup + steps):
            prep()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            torch.cuda.synchronize()
            if i >= warmup:
                ts.append(e0.elapsed_time(e1))
        return sum(ts) / len(ts)

    sampler = ClockSampler(0)
    sampler.start()
    if wl in ("sketch_sparse", "sketch_dense"):
        m, n = args.m if args.m != (1 << 24) else (1 << 23), args.n if args.n != 1024 else 2048
        dtype = torch.float32 if args.dtype == "f32" else torch.float64
        es = 4 if dtype == torch.float32 else 8
        d = args.d
        A = gen(m, n, dtype, 0xA3)
        B = rl.empty_f(d, n, dtype, dev)
        if wl == "sketch_sparse":
            D = rl.SparseDist(d, m, args.nnz)
            ms = timed(lambda: rl.sketch_general_left(ctx, D, rl.RNGState(0), A, d, 1.0, 0.0, B), lambda: None, args.steps, args.warmup)
            by = es * (m * n + 2 * d * n)                      # the reference's own model minus index bytes
            out = {"metric": "sparse_sketch_gbps", "value": by / ms / 1e6, "unit": "GB/s", "ms_per_step": ms,
                   "roofline": {"bound": "hbm", "achieved": by / ms / 1e6, "peak": hbm, "unit": "GB/s", "frac": by / ms / 1e6 / hbm,
                                "traffic": None, "peak_source": hbm_src},
                   "config": {"workload": f"SASO left sketch d={d} vec_nnz={args.nnz} of a {m} x {n} {args.dtype} matrix (configs[2] shape)"}}
        else:
            D = rl.DenseDist(d, m)
            ms = timed(lambda: rl.sketch_general_left(ctx, D, rl.RNGState(0), A, d, 1.0, 0.0, B), lambda: None, args.steps, args.warmup)
            fl = 2.0 * d * m * n
            if args.engine == "i8":
                dg = args.digits or (4 if dtype == torch.float32 else 6)
                i8_ops = fl * (dg * (dg + 1) // 2)
                b_, s_, src, _ = measured_i8_peak()
                peak = s_ or 2.0 * peaks.get("bf16_tflops_sustained", 1361.4)
                roof = {"bound": "tensor", "achieved": i8_ops / ms / 1e9, "peak": peak, "unit": "TFLOP/s", "frac": i8_ops / ms / 1e9 / peak,
                        "op": "int8 multiply-add = 2 ops; whole call (regeneration of S included in the time)", "digits": dg,
                        "traffic": None, "traffic_algorithmic": es * (m * n + d * n), "peak_source": src}
            else:
                peak, src, _ = measured_fp64_peak()
                roof = {"bound": "tensor", "achieved": fl / ms / 1e9, "peak": peak, "unit": "TFLOP/s", "frac": fl / ms / 1e9 / peak, "traffic": None,
                        "peak_source": src}
            out = {"metric": "dense_sketch_gflops", "value": fl / ms / 1e6, "unit": "Gflop/s", "ms_per_step": ms,
                   "roofline": roof,
                   "config": {"workload": f"Gaussian left sketch d={d} (operator regenerated on chip) of a {m} x {n} {args.dtype} matrix"}}
    elif wl == "cqrrpt":
        m, n = args.m if args.m != (1 << 24) else (1 << 23), args.n if args.n != 1024 else 2048
        dtype = torch.float32 if args.dtype == "f32" else torch.float64
        d_factor, nnz = args.d_factor, args.nnz
        d = int(d_factor * n)
        A = rl.empty_f(m, n, dtype, dev)
        fn = fill32 if dtype == torch.float32 else fill64
        R = torch.zeros((n, n), dtype=dtype, device=dev).t()
        J = torch.zeros(n, dtype=torch.int64, device=dev)
        alg = rl.CQRRPT(False, None)
        alg.nnz = nnz

        def prep():
            ctx.check(fn(ctx._h, m, n, rl.FAMILY_GAUSSIAN, rl.AXIS_LONG, rl.LAYOUT_NATURAL, m, n, 0, 0, A.data_ptr(), rl.RNGState(0xA3).words()))
            R.zero_()

        def step():
            rc, _, _ = alg.call(ctx, A, d_factor, rl.RNGState(0), R=R, J=J)
            assert rc == 0 and alg.rank == n, (rc, alg.rank)
        ms = timed(step, prep, args.steps, args.warmup)
        fl = 3.0 * m * n * n + m * n * nnz + (2.0 * d * n * n - 2.0 / 3.0 * n ** 3)
        if args.engine == "i8":
            # the three O(m n^2) products (A R_sk^-1, Gram, A R^-1: m n^2 flops each with the triangular / upper-only skipping) run as digit-pair
            # GEMMs on tcgen05: 4 digits (10 pairs) for fp32 storage, 7 digits (28 pairs) for fp64
            dg = args.digits or (4 if dtype == torch.float32 else 7)
            i8_ops = 3.0 * m * n * n * (dg * (dg + 1) // 2)
            b_, s_, src, _ = measured_i8_peak()
            peak = s_ or 2.0 * peaks.get("bf16_tflops_sustained", 1361.4)
            roof = {"bound": "tensor", "achieved": i8_ops / ms / 1e9, "peak": peak, "unit": "TFLOP/s", "frac": i8_ops / ms / 1e9 / peak,
                    "op": "int8 multiply-add = 2 ops; whole step (sketch, QRCP of the sketch and permutation included in the time)",
                    "digits": dg, "traffic": None, "peak_source": src}
        else:
            peak, src, _ = measured_fp64_peak()
            roof = {"bound": "tensor", "achieved": fl / ms / 1e9, "peak": peak, "unit": "TFLOP/s", "frac": fl / ms / 1e9 / peak, "traffic": None,
                    "peak_source": src}
        out = {"metric": "cqrrpt_gflops", "value": fl / ms / 1e6, "unit": "Gflop/s", "ms_per_step": ms,
               "roofline": roof,
               "config": {"workload": f"CQRRPT of a {m} x {n} {args.dtype} Gaussian matrix, SASO d={d} vec_nnz={nnz}, geqp3 (configs[2])",
                          "engine": args.engine}}
    elif wl == "bqrrp":
        n = args.n if args.n != 1024 else 65536
        m = args.m if args.m != (1 << 24) else n
        dtype = torch.float32 if args.dtype == "f32" else torch.float64
        b = args.block
        A = rl.empty_f(m, n, dtype, dev)
        fn = fill32 if dtype == torch.float32 else fill64
        tau = torch.zeros(n, dtype=dtype, device=dev)
        J = torch.zeros(n, dtype=torch.int64, device=dev)
        alg = rl.BQRRP(False, b)
        alg.qr_tall = rl.QRTALL_CHOLQR

        def prep():
            ctx.check(fn(ctx._h, max(m, n), min(m, n), rl.FAMILY_GAUSSIAN, rl.AXIS_LONG, rl.LAYOUT_COLMAJOR if m >= n else rl.LAYOUT_ROWMAJOR,
                         max(m, n), min(m, n), 0, 0, A.data_ptr(), rl.RNGState(0xA4).words()))

        def step():
            rc, _, _ = alg.call(ctx, A, args.d_factor, rl.RNGState(0), tau=tau, J=J)
            assert rc == 0 and alg.rank == min(m, n), (rc, alg.rank)
        ms = timed(step, prep, args.steps, args.warmup)
        d = int(args.d_factor * b)
        fl = 2.0 * m * n * n - 2.0 / 3.0 * n ** 3 + 2.0 * d * m * n
        if args.engine == "i8":
            # the trailing updates (the O(m n^2) part) run as 28 digit pairs (7 digits) on tcgen05; panels and the sketch on the fp64 pipe
            dg = args.digits or 7
            i8_ops = (2.0 * m * n * n - 2.0 / 3.0 * n ** 3) * (dg * (dg + 1) // 2)
            b_, s_, src, _ = measured_i8_peak()
            peak = s_ or 2.0 * peaks.get("bf16_tflops_sustained", 1361.4)
            roof = {"bound": "tensor", "achieved": i8_ops / ms / 1e9, "peak": peak, "unit": "TFLOP/s", "frac": i8_ops / ms / 1e9 / peak,
                    "op": "int8 multiply-add = 2 ops; whole step", "digits": dg, "traffic": None, "peak_source": src}
        else:
            peak, src, _ = measured_fp64_peak()
            roof = {"bound": "tensor", "achieved": fl / ms / 1e9, "peak": peak, "unit": "TFLOP/s", "frac": fl / ms / 1e9 / peak, "traffic": None,
                    "peak_source": src}
        out = {"metric": "bqrrp_gflops", "value": fl / ms / 1e6, "unit": "Gflop/s", "ms_per_step": ms,
               "roofline": roof,
               "config": {"workload": f"BQRRP of a {m} x {n} {args.dtype} Gaussian matrix, b={b}, d_factor={args.d_factor}, luqr + cholqr/orhr_col "
                                      "panels + compact-WY update (configs[3])"}}
    elif wl == "hqrrp":
        # RandLAPACK::hqrrp (rl_hqrrp.hh:811) on a tall Gaussian matrix: nb_alg = --block, pp = nb_alg / 8, CholQR + Householder-reconstruction
        # panels (qr_type 2) unless --hqrrp-panel says otherwise.  SURVEY 8(f) row 2; not one of BASELINE.json's configs.
        m = args.m if args.m != (1 << 24) else 32768
        n = args.n if args.n != 1024 else 4096
        dtype = torch.float32 if args.dtype == "f32" else torch.float64
        nb, pp = args.block if args.block != 256 else 128, max(1, (args.block if args.block != 256 else 128) // 8)
        piv, qt = {"cholqr": (0, 2), "geqrf": (0, 1), "pivoted": (1, 0)}[args.hqrrp_panel]
        A = rl.empty_f(m, n, dtype, dev)
        fn = fill32 if dtype == torch.float32 else fill64
        tau = torch.zeros(n, dtype=dtype, device=dev)
        J = torch.zeros(n, dtype=torch.int64, device=dev)

        def prep():
            ctx.check(fn(ctx._h, max(m, n), min(m, n), rl.FAMILY_GAUSSIAN, rl.AXIS_LONG, rl.LAYOUT_COLMAJOR if m >= n else rl.LAYOUT_ROWMAJOR,
                         max(m, n), min(m, n), 0, 0, A.data_ptr(), rl.RNGState(0xA5).words()))

        def step():
            rc, _, _ = rl.hqrrp(ctx, A, nb, pp, piv, qt, rl.RNGState(0), tau=tau, J=J)
            assert rc == 0, rc
        ms = timed(step, prep, args.steps, args.warmup)
        fl = 2.0 * m * n * n - 2.0 / 3.0 * n ** 3 + 2.0 * (nb + pp) * m * n
        peak, src, _ = measured_fp64_peak()
        roof = {"bound": "tensor", "achieved": fl / ms / 1e9, "peak": peak, "unit": "TFLOP/s", "frac": fl / ms / 1e9 / peak, "traffic": None,
                "peak_source": src, "note": "whole step against the fp64 pipe; the trailing updates of blocks with >= 8192 rows below them run on the int8 engine"}
        out = {"metric": "hqrrp_gflops", "value": fl / ms / 1e6, "unit": "Gflop/s", "ms_per_step": ms, "roofline": roof,
               "config": {"workload": f"hqrrp of a {m} x {n} {args.dtype} Gaussian matrix, nb_alg={nb}, pp={pp}, panels: {args.hqrrp_panel} (SURVEY 8f row 2)"}}
    else:
        raise SystemExit(f"unknown workload {wl}")
    clocks = sampler.stop()
    # per-class CUDA-event times of one extra, untimed step
    if wl in ("cqrrpt", "bqrrp", "hqrrp"):
        names = ["gemm_nn", "gemm_tn", "rightmul", "small", "fill", "sketch", "factor", "i8_mma_nn", "i8_mma_tn", "i8_slice"]
        ctx.timers_enable(True)
        for w in range(len(names)):
            ctx.timer_read(w, reset=True)
        prep()
        step()
        torch.cuda.synchronize()
        tm = {nm: ctx.timer_read(i) for i, nm in enumerate(names)}
        ctx.timers_enable(False)
        out["class_ms_per_step"] = {k_: round(v[0], 3) for k_, v in tm.items() if v[1]}
        out["class_launches_per_step"] = {k_: v[1] for k_, v in tm.items() if v[1]}
    cpu = None
    if wl in ("cqrrpt", "bqrrp", "hqrrp") and not args.no_cpu:
        # the reference's own CPU driver (oracle/_ref/librl_ref.so = its unmodified headers over OpenBLAS) on a bounded sample of the same
        # workload, all host threads; rates are size-normalised with the same flop formula
        try:
            import numpy as np
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import _ref
            from oracle import rl_oracle as O
            Rl = _ref.ref_lib()
            cores = os.cpu_count() or 1
            Rl.rlref_set_num_threads(cores)
            npdt = np.float32 if args.dtype == "f32" else np.float64
            if wl == "cqrrpt":
                mc, nc = 1 << 17, n
                Ac, _ = O.fill_dense(mc, nc, O.RNGState(0xA3), dtype=npdt)
                Ac = np.asfortranarray(Ac)
                t0 = time.perf_counter()
                _ref.ref_cqrrpt(Rl, Ac, d_factor, [0] * 6, None, nnz)
                tc = time.perf_counter() - t0
                dc = int(d_factor * nc)
                flc = 3.0 * mc * nc * nc + mc * nc * nnz + (2.0 * dc * nc * nc - 2.0 / 3.0 * nc ** 3)
                sample = f"{mc} x {nc} {args.dtype} rows of the same workload, RandLAPACK::CQRRPT (geqp3), {cores} threads, one run"
            elif wl == "hqrrp":
                mc, nc = 8192, 2048
                Ac, _ = O.fill_dense(mc, nc, O.RNGState(0xA5), dtype=npdt)
                Ac = np.asfortranarray(Ac)
                t0 = time.perf_counter()
                _ref.ref_hqrrp(Rl, Ac, nb, pp, piv, qt, [0] * 6)
                tc = time.perf_counter() - t0
                flc = 2.0 * mc * nc * nc - 2.0 / 3.0 * nc ** 3 + 2.0 * (nb + pp) * mc * nc
                sample = f"{mc} x {nc} {args.dtype}, RandLAPACK::hqrrp (nb_alg={nb}, pp={pp}, panels: {args.hqrrp_panel}), {cores} threads, one run"
            else:
                nc = 8192
                Ac, _ = O.fill_dense(nc, nc, O.RNGState(0xA4), dtype=npdt)
                Ac = np.asfortranarray(Ac)
                t0 = time.perf_counter()
                _ref.ref_bqrrp(Rl, Ac, args.d_factor, args.block, [0] * 6, 0, 1)
                tc = time.perf_counter() - t0
                flc = 2.0 * nc ** 3 - 2.0 / 3.0 * nc ** 3 + 2.0 * int(args.d_factor * args.block) * nc * nc
                sample = f"{nc} x {nc} {args.dtype}, RandLAPACK::BQRRP (luqr + cholqr, b={args.block}), {cores} threads, one run"
            cpu = {"value": flc / tc / 1e9, "unit": "Gflop/s", "cores": cores, "kind": "reference", "sample": sample, "seconds": tc}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": "Gflop/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {type(e).__name__}: {e}"}
    out.update({"n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.dtype, "data": "synthetic", "clocks": clocks, "gpu_launches": ctx.launch_count(), "cpu_baseline": cpu, "e2e": None})
    out["config"]["l2"] = "inputs exceed the 126 MB L2 by >100x; no flush needed"
    emit(out)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--m", type=int, default=1 << 24, help="rows per GPU")
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--k", type=int, default=256)
    ap.add_argument("--p", type=int, default=2, help="RS passes_over_data")
    ap.add_argument("--q", type=int, default=1, help="RS passes_per_stab")
    ap.add_argument("--m-cpu", type=int, default=1 << 17, help="rows of the bounded CPU-baseline sample")
    ap.add_argument("--m-e2e", type=int, default=1 << 20, help="rows of the host-buffer (e2e) measurement")
    ap.add_argument("--workload", default="rsvd", choices=["rsvd", "cqrrpt", "bqrrp", "hqrrp", "sketch_sparse", "sketch_dense"])
    ap.add_argument("--hqrrp-panel", default="cholqr", choices=["cholqr", "geqrf", "pivoted"], help="hqrrp workload: the panel QR")
    ap.add_argument("--dtype", default=None, help="secondary workloads: f32 | f64")
    ap.add_argument("--d", type=int, default=4096, help="sketch rows (sketch workloads)")
    ap.add_argument("--nnz", type=int, default=1, help="SASO non-zeros per column")
    ap.add_argument("--d-factor", type=float, default=None)
    ap.add_argument("--block", type=int, default=256, help="BQRRP block size")
    ap.add_argument("--engine", default="i8", choices=["dmma", "i8"],
                    help="tall fp64 products over A: tcgen05 int8 digit slices (default) or the fp64 DMMA pipe")
    ap.add_argument("--digits", type=int, default=0, help="int8 digits per value (0 = default: 6 for fp64, 46 bits)")
    ap.add_argument("--config", default="c2", choices=["c2", "c5"],
                    help="c2 (default): BASELINE.json configs[1], 2^24 x 1024 k=256 per GPU (weak scaling); c5: configs[4], 128M x 512 k=128 "
                         "row-sharded - strong scaling where 2^27 / N rows fit one GPU (N >= 4), else 2^25 rows per GPU (weak)")
    ap.add_argument("--stab", default="cholqrq", choices=["cholqrq", "plul", "hqrq"],
                    help="RS stabiliser: cholqrq (default, row-shardable, CholQR folded into the next product) or the reference's canonical "
                         "PLUL (test/drivers/test_rsvd.cc:69-93) / HQRQ; RF and QB always use CholQRQ")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    _capture_stdout()
    if args.workload != "rsvd":
        if args.dtype is None:
            args.dtype = "f64" if args.workload in ("bqrrp", "hqrrp") else "f32"
        if args.d_factor is None:
            args.d_factor = 1.0 if args.workload == "bqrrp" else 2.0
        return run_secondary(args)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    scaling = "weak"
    if args.config == "c5":
        args.n, args.k = 512, 128
        total = 1 << 27
        args.m = min(total // world, 1 << 25)
        scaling = "strong" if args.m * world == total else "weak"
    n, k, p, q = args.n, args.k, args.p, args.q
    config = {"workload": f"rank-{k} RSVD of a ({world}x{args.m}) x {n} fp64 Gaussian matrix (BASELINE.json {'configs[4], ' + scaling + ' scaling' if args.config == 'c5' else 'configs[1] per GPU'}), "
                          f"RS(p={p}, q={q}, CholQRQ) + RF(CholQRQ) + QB(block={k}, CholQRQ) + RSVD",
              "m_per_gpu": args.m, "n": n, "k": k, "passes_over_data": p, "passes_per_stab": q, "stabiliser": "CholQRQ",
              "l2": "inputs (A: m*n*8 bytes per GPU) exceed the 126 MB L2 by >1000x; no flush needed",
              "parallelism": f"row-sharded x{world}" if world > 1 else "single GPU"}

    if args.impl == "reference":
        # the reference's own CPU implementation of the path on the host cores, bounded sample of the same workload
        if rank != 0:
            return 0
        gf, t, kind, cores = cpu_sample(n, k, p, q, args.m_cpu, max(1, args.steps), max(0, min(args.warmup, 1)))
        sample = f"{args.m_cpu} x {n} fp64 rows of the same workload (k={k}, p={p}), {cores} threads"
        emit({"impl": "reference", "metric": METRIC, "value": gf, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": scaling,
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": gf, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                          "e2e": {"value": gf, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0})
        return 0

    import torch
    import randlapack_b200 as rl

    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist = None
        torch.cuda.set_device(0)
    dev = torch.device("cuda", torch.cuda.current_device())
    ctx = rl.Context(dev.index)
    ctx.set_fp64_engine(args.engine)
    ctx.set_i8_digits(args.digits)
    config["fp64_engine"] = ("tcgen05 kind::i8 digit slices fused with the slicing of A, %d digits (A^T*Y in the power iteration, Gram matrix and U: %d)"
                             % (args.digits or 6, args.digits or 7)) if args.engine == "i8" else "DMMA fp64 pipe"

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic resident input: A = Gaussian DenseDist(m_global, n) sample, generated ON DEVICE by our own fill kernel
    m_local = args.m
    free_b, total_b = torch.cuda.mem_get_info()
    need = 8 * (m_local * n + m_local * k) + (4 << 30)
    while need > free_b and m_local > (1 << 16):
        m_local //= 2
        need = 8 * (m_local * n + m_local * k) + (4 << 30)
    if m_local != args.m:
        config["workload"] += f" [REDUCED to {m_local} rows per GPU: only {free_b / 1e9:.1f} GB free]"
        config["m_per_gpu"] = m_local
    m_global = m_local * world
    A = rl.empty_f(m_local, n, torch.float64, dev)
    buf, _ = None, None
    st_in = rl.RNGState(0xA2)
    fn = ctx._lib.rlb200_fill_dense_f64_dev
    ctx.check(fn(ctx._h, m_global, n, rl.FAMILY_GAUSSIAN, rl.AXIS_LONG, rl.LAYOUT_NATURAL, m_local, n, rank * m_local, 0,
                 A.data_ptr(), st_in.words()))
    U = rl.empty_f(m_local, k, torch.float64, dev)
    S = torch.empty(k, dtype=torch.float64, device=dev)
    V = rl.empty_f(n, k, torch.float64, dev)
    if world > 1:
        ctx.set_shard(rank * m_local, m_global)
    stab = {"cholqrq": rl.CholQRQ, "plul": rl.PLUL, "hqrq": rl.HQRQ}[args.stab]()
    stack = rl.RSVD(rl.QB(rl.RF(rl.RS(stab, p, q), rl.CholQRQ()), rl.CholQRQ()), k)
    config["stabiliser"] = {"cholqrq": "CholQRQ", "plul": "PLUL (RS) + CholQRQ (RF, QB)", "hqrq": "HQRQ (RS) + CholQRQ (RF, QB)"}[args.stab]

    def step():
        rc, kk, *_ = stack.call(ctx, A, k, 0.0, rl.RNGState(0), U=U, S=S, V=V)
        assert rc == 0 and kk == k, (rc, kk, stack.qb_code)

    TIMER_NAMES = ["gemm_nn", "gemm_tn", "rightmul", "small", "fill", "sketch", "factor", "i8_mma_nn", "i8_mma_tn", "i8_slice"]
    for _ in range(args.warmup):
        step()
    barrier()
    ctx.launch_count(reset=True)
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = int(round(ctx.launch_count() / args.steps))
    # per-class CUDA-event times on the launching streams: one extra, untimed step (the class timers synchronise the host between
    # launches, so they are kept out of the timed region)
    ctx.timers_enable(True)
    for w in range(len(TIMER_NAMES)):
        ctx.timer_read(w, reset=True)
    step()
    torch.cuda.synchronize()
    tms = {nm: ctx.timer_read(i) for i, nm in enumerate(TIMER_NAMES)}
    ctx.timers_enable(False)
    if dist is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = t.item()
    ms_step = ms_total / args.steps
    value = rsvd_flops(m_global, n, k, p, q) / (ms_step * 1e-3) / 1e9

    # ---- sanity of the timed result (not timed): orthonormal U on a row sample is meaningless when sharded; check sigma > 0, finite
    assert torch.isfinite(S).all() and (S[:-1] >= S[1:]).all() and S[-1] > 0

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (per-class CUDA-event times were recorded on the launching stream)
    cf = class_flops(m_local, n, k, p, q)
    peaks = load_peaks()
    class_ms = {kname: v[0] for kname, v in tms.items() if v[1]}
    class_launches = {kname: v[1] for kname, v in tms.items() if v[1]}
    fp64_peak, fp64_src, _ = measured_fp64_peak()
    if args.engine == "i8":
        # dominant kernel: oz2_kernel (tcgen05.mma.kind::i8 with the digit slicing of A fused in).  Work per launch class: the (p + 2) tall
        # products, 2*m*n*k flops each, executed as digit-pair GEMMs of the same shape:
        #   A*Omega passes: 6 digits, 21 pairs;  U = Y*M: 7 digits, 28 pairs (orthogonality of U);
        #   A^T*Y inside the power iteration (Omega = (A^T Y) R^-1): 7 digits, 28 pairs;  the last A^T*Y (B^T): 21 pairs;
        #   the Gram tiles of Y^T*Y that ride in every A^T*Y launch: 28 pairs on the tiles that touch the upper triangle.
        S_dig = args.digits or 6
        S_tn = args.digits or 7
        pairs, pairs_full = S_dig * (S_dig + 1) // 2, S_tn * (S_tn + 1) // 2
        nn_passes = p // 2 + 1
        tn_fold = p // 2                      # A^T*Y launches inside RS that carry the folded CholQR (full pairs)
        tn_plain = 1 if p % 2 else 0          # odd p: the first A^T*Omega_1 (no Gram matrix)
        gram_rows, gram_cols = -(-k // 128), -(-k // 64)
        gram_frac = sum(1 for y in range(gram_rows) for x in range(gram_cols) if x * 64 <= y * 128 + 127) / float(gram_rows * gram_cols)
        mma_ms = tms["i8_mma_nn"][0] + tms["i8_mma_tn"][0]
        i8_ops = nn_passes * 2.0 * m_local * n * k * pairs + 2.0 * m_local * k * k * pairs_full \
            + tn_fold * (2.0 * m_local * n * k + 2.0 * m_local * k * k * gram_frac) * pairs_full \
            + tn_plain * 2.0 * m_local * n * k * pairs \
            + (2.0 * m_local * n * k * pairs + 2.0 * m_local * k * k * gram_frac * pairs_full)
        achieved = i8_ops / (mma_ms * 1e-3) / 1e12 if mma_ms > 0 else 0.0
        i8_burst, i8_sus, i8_src, i8_detail = measured_i8_peak()
        if i8_sus:
            peak, peak_src = i8_sus, i8_src + "; sustained figure (the kernel is timed inside a long step); burst %.0f" % i8_burst
        elif "bf16_tflops_sustained" in peaks:
            peak = 2.0 * peaks["bf16_tflops_sustained"]
            peak_src = "2 x MEASURED_PEAKS.json bf16_tflops_sustained (tools/peaks_i8 did not run: " + i8_src + ")"
        else:
            peak = 2.0 * 1361.4
            peak_src = "fallback: 2 x 1361.4 TFLOP/s sustained bf16 (B200_PROFILING.md)"
        fused_traffic = (m_local * (OZ2_TRAFFIC_NN_PER_ROW * nn_passes + OZ2_TRAFFIC_TN_PER_ROW * (tn_fold + tn_plain + 1))
                         if (k == 256 and n == 1024 and S_dig == 6) else None)
        roofline = {"bound": "tensor", "kernel": "oz2_kernel (tcgen05.mma.kind::i8 + fused digit slicing of A, NN + TN launches)", "achieved": achieved,
                    "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "op": "int8 multiply-add = 2 ops",
                    "whole_step_frac": i8_ops / (ms_step * 1e-3) / 1e12 / peak,
                    "fp64_equivalent_tflops": ((p + 2) * 2.0 * m_local * n * k + 2.0 * m_local * k * k * (1 + (tn_fold + 1) * gram_frac))
                    / (mma_ms * 1e-3) / 1e12 if mma_ms > 0 else None,
                    "binding_resource": "L1 data pipe of the SM (tensor-core operand reads 46 % + shared-memory / global LSU traffic 34 % of its peak in "
                                        "the A*Omega launch: ncu --set full, profiles/ncu_oz3_nn_r2i.txt; DESIGN.md 3b) - not HBM (37 %) and not the tensor pipe",
                    "digits": {"a_omega": S_dig, "u_and_at_y_in_power_iteration_and_gram": S_tn},
                    "digit_pairs": {"a_omega_and_last_at_y": pairs, "u_and_at_y_in_power_iteration_and_gram": pairs_full},
                    "int8_ops_per_step": i8_ops,
                    "traffic": (STEP_TRAFFIC_PER_ROW * m_local) if (STEP_TRAFFIC_PER_ROW and k == 256 and n == 1024 and p == 2) else fused_traffic,
                    "traffic_dominant_launches_only": fused_traffic,
                    "traffic_algorithmic": 8.0 * m_local * n * (p + 2),
                    "traffic_source": OZ_TRAFFIC_SOURCE, "peak_source": peak_src, "int8_peak_detail": i8_detail,
                    "class_ms_per_step": class_ms, "class_launches_per_step": class_launches,
                    "fp64_pipe_peak_tflops": fp64_peak, "fp64_pipe_peak_source": fp64_src,
                    "whole_step_vs_fp64_pipe_peak": (value / 1e3) / fp64_peak if fp64_peak else None}
    else:
        nn_ms, tn_ms = tms["gemm_nn"][0], tms["gemm_tn"][0]
        dom = "gemm_nn" if nn_ms >= tn_ms else "gemm_tn"
        dom_ms = max(nn_ms, tn_ms)
        dom_flops = (cf["gemm_nn"] + cf["rightmul"]) if dom == "gemm_nn" else cf["gemm_tn"]
        achieved = dom_flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
        # DRAM traffic of the dominant kernel per launch: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture
        # of gemm_nn_kernel<128,64,...> at m = 2^21 (profiles/ncu_gemm_nn_r1.txt: 17.193 GB + 4.280 GB for 17.180 + 4.295 GB of
        # algorithmic bytes, i.e. A and Y each cross HBM exactly once), scaled by m; ~the same per launch for A^T*Y (reads A and Y).
        traffic = (17.193092e9 + 4.279688e9) * (m_local / float(1 << 21)) * (n / 1024.0) if k == 256 else None
        roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": achieved / fp64_peak if fp64_peak else None, "traffic": traffic,
                    "traffic_source": "ncu --set full capture at m=2^21 scaled by m (profiles/ncu_gemm_nn_r1.txt)",
                    "peak_source": fp64_src + "; fp64 pipe (MEASURED_PEAKS.json has no fp64 figure)",
                    "class_ms_per_step": class_ms, "class_launches_per_step": class_launches,
                    "whole_step_frac_of_fp64_peak": (value / 1e3) / fp64_peak if fp64_peak else None}

    # ---- e2e: the reference-facing call with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e and world == 1:
        try:
            m_e = args.m_e2e
            avail_kb = [int(l.split()[1]) for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
            while 8 * m_e * n * 2.2 > avail_kb * 1024 * 0.5 and m_e > (1 << 14):
                m_e //= 2
            del A, U
            torch.cuda.empty_cache()
            Ah = torch.empty((n, m_e), dtype=torch.float64, pin_memory=True).t()
            tmp = rl.empty_f(m_e, n, torch.float64, dev)
            ctx.check(fn(ctx._h, m_e, n, rl.FAMILY_GAUSSIAN, rl.AXIS_LONG, rl.LAYOUT_NATURAL, m_e, n, 0, 0, tmp.data_ptr(),
                         rl.RNGState(0xA2).words()))
            Ah.copy_(tmp)
            del tmp
            torch.cuda.synchronize()
            Uh = torch.empty((k, m_e), dtype=torch.float64, pin_memory=True).t()
            Sh = torch.empty(k, dtype=torch.float64, pin_memory=True)
            Vh = torch.empty((k, n), dtype=torch.float64, pin_memory=True).t()
            for _ in range(2):
                stack.call_host(ctx, Ah, k, 0.0, rl.RNGState(0), U=Uh, S=Sh, V=Vh)
            t0 = time.perf_counter()
            reps = max(2, args.steps)
            for _ in range(reps):
                rc, kk, *_ = stack.call_host(ctx, Ah, k, 0.0, rl.RNGState(0), U=Uh, S=Sh, V=Vh)
            te = (time.perf_counter() - t0) / reps
            e2e = {"value": rsvd_flops(m_e, n, k, p, q) / te / 1e9, "unit": UNIT, "h2d_bytes_per_step": 8 * m_e * n,
                   "d2h_bytes_per_step": 8 * (m_e * k + k + n * k), "ms_per_step": te * 1e3,
                   "workload": f"{m_e} x {n} fp64 host-resident A (pinned in, pinned out), rlb200_rsvd_f64_host: H2D + RSVD + D2H of U,S,V"}
        except Exception as e:  # noqa: BLE001
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": f"{type(e).__name__}: {e}"}

    parity = None
    if not args.no_cpu and world == 1:
        try:
            parity = parity_sample(ctx, rl, n, k, p, q, min(args.m_cpu, 1 << 17))
        except Exception as e:  # noqa: BLE001
            parity = {"error": f"{type(e).__name__}: {e}"}
    cpu = None
    if not args.no_cpu and world == 1:
        try:
            gf, t, kind, cores = cpu_sample(n, k, p, q, args.m_cpu, 2, 1)
            cpu = {"value": gf, "unit": UNIT, "cores": cores, "kind": kind, "ms_per_step": t * 1e3,
                   "sample": f"{args.m_cpu} x {n} fp64 rows of the same workload (k={k}, p={p}), 1 warm-up + 2 timed runs"}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {type(e).__name__}: {e}"}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": config, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "parity": parity,
           "gpu_launches": launches, "flops_per_step": rsvd_flops(m_global, n, k, p, q),
           "flops_executed_per_step": sum(class_flops(m_global, n, k, p, q).values()),
           "note": "value = nominal algorithm flops F(m,n,k,p) (DESIGN.md, same F as the reference arm) / time; CholQR's m x k "
                   "triangular solves are folded into the next product and not executed (flops_executed_per_step)"}
    emit(out)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
