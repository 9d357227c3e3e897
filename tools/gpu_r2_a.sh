#!/bin/bash
# round 2, step A: measured int8 peak + instruction-mix costs; fused engine (LDG -> registers -> digits) correctness and timing
mkdir -p gpurun_out
timeout 180 tools/peaks_i8 2 > gpurun_out/peaks_i8_r2a.json 2> gpurun_out/peaks_i8_r2a.err; echo "peaks rc=$?"; cat gpurun_out/peaks_i8_r2a.json | head -c 3000; echo; tail -3 gpurun_out/peaks_i8_r2a.err
python -m pytest tests/test_gpu_ozaki.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_r2a_oz.log; tail -4 gpurun_out/pytest_r2a_oz.log
for ng in 3 2; do
  echo "== NG=$ng dbg=1"
  RLB200_OZ2_NG=$ng RLB200_OZ2_DBG=1 RLB200_OZ_ASSUME_CONST=1 timeout 300 python tools/bench_gemm.py 21 1024 256 i8 > /tmp/o.txt 2>&1
  grep "oz2 dbg\] NN" /tmp/o.txt | head -1 | cut -c1-330; grep "oz2 dbg\] TN" /tmp/o.txt | head -1 | cut -c1-330
  echo "== NG=$ng"
  RLB200_OZ2_NG=$ng RLB200_OZ_ASSUME_CONST=1 timeout 300 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | cut -c1-300 | tail -2
done
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_r2a_all.log; tail -8 gpurun_out/pytest_r2a_all.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
python -c "import json; d=json.load(open('gpurun_out/bench_r2a.json')); print('rsvd', d['ms_per_step'], d['value'], d['roofline'].get('class_ms_per_step'))"
tail -3 gpurun_out/bench_r2a.err
