#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cqrrpt.py tests/test_gpu_bqrrp.py tests/test_gpu_ozaki.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_r2j.log; tail -12 gpurun_out/pytest_r2j.log | cut -c1-400
# ncu of the current NN kernel (2 converter groups, digits not shared)
RLB200_OZ_ASSUME_CONST=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:oz2_kernel --launch-skip 1 -c 1 -f -o gpurun_out/prof_oz2_nn_r2g python tools/bench_gemm.py 20 1024 256 nn > gpurun_out/ncu_oz2_nn_r2g.log 2>&1
tail -1 gpurun_out/ncu_oz2_nn_r2g.log
