#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bqrrp.py tests/test_gpu_drivers.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_bq.log
tail -3 gpurun_out/pytest_bq.log
timeout 300 python bench.py --workload bqrrp --n 32768 --steps 1 --warmup 1 > gpurun_out/sec_bqrrp_32768_i8.json 2> gpurun_out/sec.err; tail -2 gpurun_out/sec.err
python - <<PY
import json; d=json.load(open("gpurun_out/sec_bqrrp_32768_i8.json")); print(d["ms_per_step"], d["value"], d.get("class_ms_per_step"), d.get("class_launches_per_step"))
PY
