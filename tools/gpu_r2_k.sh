#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -2
for ng in 2 3; do echo "== NG=$ng"; RLB200_OZ2_NG=$ng RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | cut -c1-250 | tail -1; done
echo "== NG=2 writer fence"; RLB200_OZ2_DBG=128 RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | cut -c1-250 | tail -1
RLB200_OZ2_DBG=1 RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 > /tmp/o.txt 2>&1
grep "oz2 dbg\] NN" /tmp/o.txt | head -1 | cut -c1-330; grep "oz2 dbg\] TN" /tmp/o.txt | head -1 | cut -c1-330
timeout 400 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_cqrrpt.py -m gpu -q 2>&1 | tail -3
