#!/bin/bash
for ns in 0 1; do
  if [ $ns = 1 ]; then export RLB200_OZ2_NOSHARE=1; else unset RLB200_OZ2_NOSHARE; fi
  echo "== noshare=$ns dbg=1"
  RLB200_OZ2_DBG=1 RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 > /tmp/o.txt 2>&1
  grep "oz2 dbg\] NN" /tmp/o.txt | head -1 | cut -c1-330; grep "oz2 dbg\] TN" /tmp/o.txt | head -1 | cut -c1-330
  echo "== noshare=$ns"
  RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | cut -c1-250 | tail -1
  RLB200_OZ2_NG=2 RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | cut -c1-250 | tail -1
done
timeout 120 python -m pytest tests/test_gpu_ozaki.py -m gpu -q -x 2>&1 | tail -3
