#!/bin/bash
# ncu --set full on one NN and one TN launch of the fused kernel (m = 2^20 to keep the capture short)
mkdir -p gpurun_out
RLB200_OZ_ASSUME_CONST=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz2_kernel -c 2 -f -o gpurun_out/prof_oz2_r2e python tools/bench_gemm.py 20 1024 256 i8 > gpurun_out/ncu_oz2_r2e.log 2>&1
tail -3 gpurun_out/ncu_oz2_r2e.log
ls -la gpurun_out/prof_oz2_r2e.ncu-rep
