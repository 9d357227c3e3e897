#!/bin/bash
# fused engine: correctness, then cycle breakdown of the converter / issuer and timing experiments
# (RLB200_OZ2_DBG bits: 1 stamps, 2 no raw loads, 4 no fence, 8 no math)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ozaki.py -m gpu -q -x -k fused 2>&1 | tail -5
for ng in 3 2; do
for f in 1 3; do
  echo "== RLB200_OZ2_NG=$ng RLB200_OZ2_DBG=$f"
  RLB200_OZ2_NG=$ng RLB200_OZ2_DBG=$f RLB200_OZ_ASSUME_CONST=1 timeout 300 python tools/bench_gemm.py 21 1024 256 i8 > /tmp/o.txt 2>&1
  grep "oz2 dbg\] NN" /tmp/o.txt | head -1 | cut -c1-330; grep "oz2 dbg\] TN" /tmp/o.txt | head -1 | cut -c1-330; grep '"m"' /tmp/o.txt | cut -c1-300
done
echo "== NG=$ng no dbg"
RLB200_OZ2_NG=$ng RLB200_OZ_ASSUME_CONST=1 timeout 300 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | cut -c1-300 | tail -2
done
