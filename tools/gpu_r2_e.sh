#!/bin/bash
mkdir -p gpurun_out
RLB200_OZ_ASSUME_CONST=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:oz3_kernel --launch-skip 1 -c 1 -f -o gpurun_out/prof_oz3_nn_r2h python tools/bench_gemm.py 20 1024 256 nn > gpurun_out/ncu_oz3_nn_r2h.log 2>&1
tail -1 gpurun_out/ncu_oz3_nn_r2h.log
