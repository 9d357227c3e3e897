#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_ozaki.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_r2f_oz.log; tail -3 gpurun_out/pytest_r2f_oz.log
echo "== NG=3 dbg=1"
RLB200_OZ2_DBG=1 RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 > /tmp/o.txt 2>&1
grep "oz2 dbg\] NN" /tmp/o.txt | head -1 | cut -c1-330; grep "oz2 dbg\] TN" /tmp/o.txt | head -1 | cut -c1-330
echo "== NG=3"
RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | cut -c1-300 | tail -1
echo "== NG=3 plain stores"
RLB200_OZ2_DBG=16 RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | cut -c1-300 | tail -1
echo "== NG=2"; RLB200_OZ2_NG=2 RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | cut -c1-300 | tail -1
echo "== NG=3 noshare"; RLB200_OZ2_NOSHARE=1 RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | cut -c1-300 | tail -1
echo "== k=512 / k=128"
RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 512 i8 2>&1 | cut -c1-300 | tail -1
RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 128 i8 2>&1 | cut -c1-300 | tail -1
