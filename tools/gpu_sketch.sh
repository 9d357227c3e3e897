#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sketch.py tests/test_gpu_cqrrpt.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_sketch.log
tail -3 gpurun_out/pytest_sketch.log
python bench.py --workload sketch_sparse --nnz 1 --steps 5 --warmup 3 > gpurun_out/sec_sketch_sparse_nnz1.json 2> gpurun_out/sec.err; tail -2 gpurun_out/sec.err; cut -c1-330 gpurun_out/sec_sketch_sparse_nnz1.json
python bench.py --workload sketch_sparse --nnz 4 --steps 5 --warmup 3 > gpurun_out/sec_sketch_sparse_nnz4.json 2> gpurun_out/sec.err; tail -2 gpurun_out/sec.err; cut -c1-330 gpurun_out/sec_sketch_sparse_nnz4.json
python bench.py --workload sketch_sparse --nnz 1 --dtype f64 --steps 5 --warmup 3 > gpurun_out/sec_sketch_sparse_f64.json 2> gpurun_out/sec.err; tail -2 gpurun_out/sec.err; cut -c1-330 gpurun_out/sec_sketch_sparse_f64.json
