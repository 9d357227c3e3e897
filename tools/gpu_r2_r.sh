#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload cqrrpt --dtype f32 --steps 2 --warmup 3 --no-cpu > gpurun_out/sec_cqrrpt_c3_r2b.json 2> gpurun_out/sec_cqrrpt_c3_r2b.err
python -c "import json; d=json.load(open('gpurun_out/sec_cqrrpt_c3_r2b.json')); print('cqrrpt c3', d['ms_per_step'], d.get('class_ms_per_step'))"
timeout 600 python bench.py --workload bqrrp --dtype f64 --n 32768 --steps 1 --warmup 1 --no-cpu > gpurun_out/sec_bqrrp_32k_r2b.json 2> gpurun_out/sec_bqrrp_32k_r2b.err
python -c "import json; d=json.load(open('gpurun_out/sec_bqrrp_32k_r2b.json')); print('bqrrp 32k', d['ms_per_step'], d.get('class_ms_per_step'))"
timeout 900 python -m pytest tests/test_gpu_cqrrpt.py tests/test_gpu_bqrrp.py tests/test_gpu_drivers.py tests/test_gpu_evd.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-300
