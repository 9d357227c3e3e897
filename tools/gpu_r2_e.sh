#!/bin/bash
# ncu --set full on one NN and one TN launch of the fused kernel (m = 2^20 to keep the capture short)
mkdir -p gpurun_out
for op in nn tn; do
RLB200_OZ_ASSUME_CONST=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:oz2_kernel --launch-skip 1 -c 1 -f -o gpurun_out/prof_oz2_${op}_r2f python tools/bench_gemm.py 20 1024 256 $op > gpurun_out/ncu_oz2_${op}_r2f.log 2>&1
tail -2 gpurun_out/ncu_oz2_${op}_r2f.log
done
ls -la gpurun_out/*.ncu-rep | tail -3
