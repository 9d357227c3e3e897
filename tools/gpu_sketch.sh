#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_cqrrpt.py -m gpu -q -x 2>&1 | tail -12 > gpurun_out/pytest_sketch.log
tail -3 gpurun_out/pytest_sketch.log
for nnz in 1 2 4; do
timeout 300 python bench.py --workload sketch_sparse --nnz $nnz --steps 5 --warmup 3 > gpurun_out/sec_sketch_sparse_nnz$nnz.json 2> gpurun_out/sec.err; tail -2 gpurun_out/sec.err; cut -c1-230 gpurun_out/sec_sketch_sparse_nnz$nnz.json
done
timeout 300 python bench.py --workload sketch_sparse --nnz 1 --dtype f64 --steps 5 --warmup 3 > gpurun_out/sec_sketch_sparse_f64.json 2> gpurun_out/sec.err; tail -2 gpurun_out/sec.err; cut -c1-230 gpurun_out/sec_sketch_sparse_f64.json
ncu --set full --clock-control none --import-source on -k regex:saso_strip_kernel -s 1 -c 1 -o gpurun_out/prof_saso -f \
    python bench.py --workload sketch_sparse --nnz 1 --m 1048576 --steps 1 --warmup 1 > gpurun_out/ncu_saso.log 2>&1
