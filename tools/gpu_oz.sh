#!/bin/bash
# int8-slice engine: parity tests + product timings on both engines, per cluster size
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ozaki.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_oz.log
tail -5 gpurun_out/pytest_oz.log
for cs in 1 2 4; do
  RLB200_OZ_CLUSTER=$cs timeout 600 python tools/bench_gemm.py 22 1024 256 > gpurun_out/bench_gemm_22_cs$cs.json 2> gpurun_out/bench_gemm.err
  echo "cs=$cs"; cat gpurun_out/bench_gemm_22_cs$cs.json; tail -3 gpurun_out/bench_gemm.err
done
