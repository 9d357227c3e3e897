#!/bin/bash
# last visit of the round: the whole GPU suite, smoke(), the headline bench with every field, p = 0, the reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -2 gpurun_out/bench_1gpu.err
timeout 300 python bench.py --steps 3 --warmup 3 --p 0 --no-cpu --no-e2e > gpurun_out/bench_1gpu_p0.json 2>> gpurun_out/bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 python bench.py --workload cqrrpt --steps 2 --warmup 1 > gpurun_out/sec_cqrrpt_i8.json 2> gpurun_out/sec.err
python - <<PY
import json
for f in ["bench_1gpu","bench_1gpu_p0","bench_ref","sec_cqrrpt_i8"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); r=d.get("roofline") or {}
        print(f, round(d["ms_per_step"],1), round(d["value"]), r.get("frac"), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
