#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_drivers.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -2
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -2 gpurun_out/bench_1gpu.err
timeout 300 python bench.py --steps 3 --warmup 3 --p 0 --no-cpu --no-e2e > gpurun_out/bench_1gpu_p0.json 2>> gpurun_out/bench_1gpu.err
python - <<PY
import json
for f in ["bench_1gpu","bench_1gpu_p0"]:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); r=d.get("roofline") or {}
    print(f, round(d["ms_per_step"],1), round(d["value"]), r.get("frac"), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"), d.get("gpu_launches"))
PY
