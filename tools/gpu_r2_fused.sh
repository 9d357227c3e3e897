#!/bin/bash
# fused-engine correctness, then timing of the headline step
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ozaki.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_r2b_oz.log
tail -5 gpurun_out/pytest_r2b_oz.log
python -m pytest tests/test_gpu_drivers.py tests/test_gpu_fill.py -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_r2b_drv.log
tail -8 gpurun_out/pytest_r2b_drv.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err
python -c "import json; d=json.load(open('gpurun_out/bench_r2b.json')); print('rsvd', d['ms_per_step'], d['value'], d['roofline'].get('class_ms_per_step'))"
tail -3 gpurun_out/bench_r2b.err
