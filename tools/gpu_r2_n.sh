#!/bin/bash
for rep in 1 2; do
echo "== two-stage rounds"; RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | grep -a "i8s6" | tail -1 | cut -c1-300
echo "== one stage"; RLB200_OZ3_ONE_STAGE=1 RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | grep -a "i8s6" | tail -1 | cut -c1-300
done
timeout 300 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -2 | cut -c1-200
