#!/bin/bash
mkdir -p gpurun_out
timeout 60 python tools/dbg_share.py 70001 1031 129 6 2>&1 | tail -3 | cut -c1-300
timeout 60 python tools/dbg_share.py 20000 1024 256 7 2>&1 | tail -3 | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-300
echo "== oz3"; RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | cut -c1-250 | tail -1
echo "== oz3 again"; RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 nn 2>&1 | cut -c1-250 | tail -1
echo "== oz3 k=128"; RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 512 128 i8 2>&1 | cut -c1-250 | tail -1
echo "== oz3 no loads"; RLB200_OZ2_DBG=66 RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 nn 2>&1 | cut -c1-250 | tail -1
