#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bqrrp.py tests/test_gpu_drivers.py tests/test_gpu_cqrrpt.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_bq.log
tail -3 gpurun_out/pytest_bq.log
timeout 300 python bench.py --workload bqrrp --n 32768 --steps 1 --warmup 1 > gpurun_out/sec_bqrrp_32768_i8.json 2> gpurun_out/sec.err; tail -2 gpurun_out/sec.err
python - <<PY
import json; d=json.load(open("gpurun_out/sec_bqrrp_32768_i8.json")); print(d["ms_per_step"], d["value"], d.get("class_ms_per_step"))
PY
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rsvd', d['ms_per_step'], d['value'], d['roofline']['class_ms_per_step'].get('small'))"
