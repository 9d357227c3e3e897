#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sketch.py -m gpu -q -x 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:saso_strip_kernel -s 1 -c 1 -o gpurun_out/prof_saso -f \
    python bench.py --workload sketch_sparse --nnz 1 --m 1048576 --steps 1 --warmup 1 > gpurun_out/ncu_saso.log 2>&1
tail -3 gpurun_out/ncu_saso.log
