#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_bqrrp.py tests/test_gpu_drivers.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_bq.log
tail -3 gpurun_out/pytest_bq.log
for eng in i8 dmma; do
python bench.py --workload bqrrp --n 32768 --engine $eng --steps 1 --warmup 1 > gpurun_out/sec_bqrrp_32768_$eng.json 2> gpurun_out/sec.err; tail -2 gpurun_out/sec.err
python - <<PY
import json; d=json.load(open("gpurun_out/sec_bqrrp_32768_$eng.json")); print("$eng", d["ms_per_step"], d["value"], d.get("class_ms_per_step"), d.get("class_launches_per_step"))
PY
done
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_quick.json 2> gpurun_out/sec.err; tail -2 gpurun_out/sec.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench_quick.json")); print("rsvd", d["ms_per_step"], d["value"], d["roofline"]["class_ms_per_step"])
PY
