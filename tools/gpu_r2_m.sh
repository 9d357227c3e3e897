#!/bin/bash
mkdir -p gpurun_out
timeout 60 python tools/dbg_share.py 70001 1031 129 6 2>&1 | tail -3 | cut -c1-300
timeout 60 python tools/dbg_share.py 20000 1024 256 7 2>&1 | tail -3 | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-300
echo "== oz3"; RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | cut -c1-250 | tail -1
echo "== oz3 k=128"; RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 512 128 i8 2>&1 | cut -c1-250 | tail -1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err
wc -l gpurun_out/bench_r2m.json
python -c "import json; d=json.load(open('gpurun_out/bench_r2m.json')); print('rsvd', d['ms_per_step'], d['value']/1e3, d['roofline']['frac'], d['roofline']['class_ms_per_step'])"
