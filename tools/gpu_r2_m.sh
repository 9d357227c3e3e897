#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-300
echo "== oz3"; RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | cut -c1-250 | tail -1
echo "== oz2"; RLB200_OZ3=0 RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | cut -c1-250 | tail -1
RLB200_OZ3=0 RLB200_OZ2_DBG=1 RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 nn > /tmp/o.txt 2>&1
grep "oz2 dbg\] NN" /tmp/o.txt | head -1 | cut -c1-330
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5 | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err
python -c "import json; d=json.load(open('gpurun_out/bench_r2m.json')); print('rsvd', d['ms_per_step'], d['value']/1e3, d['roofline']['frac'], d['roofline']['class_ms_per_step'])"
