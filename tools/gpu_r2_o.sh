#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_bqrrp.py tests/test_gpu_cqrrpt.py -m gpu -q 2>&1 | tail -4 | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r2o.json 2> gpurun_out/bench_r2o.err
python -c "import json; d=json.load(open('gpurun_out/bench_r2o.json')); print('rsvd', d['ms_per_step'], d['value']/1e3, d['roofline']['frac'], d['roofline']['class_ms_per_step'])"; tail -2 gpurun_out/bench_r2o.err
timeout 900 python bench.py --workload bqrrp --n 32768 --steps 1 --warmup 1 --no-cpu > gpurun_out/sec_bqrrp_32k_r2b.json 2> gpurun_out/sec_bq.err; python -c "
import json; d=json.load(open('gpurun_out/sec_bqrrp_32k_r2b.json')); print('bqrrp 32k (NN on DMMA)', d['ms_per_step'], d['value']/1e3, d.get('class_ms_per_step'))"
RLB200_BQRRP_NN_I8=1 timeout 900 python bench.py --workload bqrrp --n 32768 --steps 1 --warmup 1 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bqrrp 32k (NN on i8)', d['ms_per_step'], d['value']/1e3, d.get('class_ms_per_step'))"
