#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | grep -a "i8s6" | tail -1 | cut -c1-300; done
RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 512 128 i8 2>&1 | grep -a "i8s6" | tail -1 | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_gemm.py tests/test_gpu_drivers.py -m gpu -q -x 2>&1 | tail -2 | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:oz3_kernel -s 2 -c 1 --csv --log-file gpurun_out/oz3_dram.csv python tools/bench_gemm.py 21 1024 256 i8 > /dev/null 2>&1
grep -a "oz3" gpurun_out/oz3_dram.csv | cut -d, -f5,13- | cut -c1-200
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r2t.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/bench_r2t.json')); print('rsvd', d['ms_per_step'], d['value']/1e3, d['roofline']['frac'], d['roofline']['class_ms_per_step'])"
