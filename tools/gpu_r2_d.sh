#!/bin/bash
# fused + staged engines with the warp-uniform issuer: correctness, cycle breakdown, timing, full GPU test-suite, headline debug, bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ozaki.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_r2d_oz.log; tail -3 gpurun_out/pytest_r2d_oz.log
for ng in 3; do
  echo "== NG=$ng dbg=1"
  RLB200_OZ2_NG=$ng RLB200_OZ2_DBG=1 RLB200_OZ_ASSUME_CONST=1 timeout 300 python tools/bench_gemm.py 21 1024 256 i8 > /tmp/o.txt 2>&1
  grep "oz2 dbg\] NN" /tmp/o.txt | head -1 | cut -c1-330; grep "oz2 dbg\] TN" /tmp/o.txt | head -1 | cut -c1-330
  echo "== NG=$ng dbg=3 (no loads)"
  RLB200_OZ2_NG=$ng RLB200_OZ2_DBG=3 RLB200_OZ_ASSUME_CONST=1 timeout 300 python tools/bench_gemm.py 21 1024 256 i8 > /tmp/o.txt 2>&1
  grep "oz2 dbg\] NN" /tmp/o.txt | head -1 | cut -c1-330; grep "oz2 dbg\] TN" /tmp/o.txt | head -1 | cut -c1-330; grep '"m"' /tmp/o.txt | cut -c1-300
  echo "== NG=$ng"
  RLB200_OZ2_NG=$ng RLB200_OZ_ASSUME_CONST=1 timeout 300 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | cut -c1-300 | tail -2
done
echo "== NG=2"; RLB200_OZ2_NG=2 RLB200_OZ_ASSUME_CONST=1 timeout 300 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | cut -c1-300 | tail -1
echo "== staged"; RLB200_I8_FUSED=0 RLB200_OZ_ASSUME_CONST=1 timeout 300 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | cut -c1-300 | tail -1
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_r2d_all.log; tail -8 gpurun_out/pytest_r2d_all.log
timeout 600 python tools/dbg_headline.py 2>&1 | grep "gramfusion=True" | grep "fused=True" | tee gpurun_out/dbg_headline_r2d.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err
python -c "import json; d=json.load(open('gpurun_out/bench_r2d.json')); print('rsvd', d['ms_per_step'], d['value'], d['roofline'].get('class_ms_per_step'))"
tail -3 gpurun_out/bench_r2d.err
