#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_bqrrp.py -m gpu -q -x 2>&1 | tail -4
for eng in i8 dmma; do
timeout 300 python bench.py --workload sketch_dense --d 256 --m 4194304 --dtype f64 --engine $eng --steps 3 --warmup 2 > gpurun_out/sec_sketch_dense_$eng.json 2> gpurun_out/sec.err; tail -2 gpurun_out/sec.err; cut -c1-260 gpurun_out/sec_sketch_dense_$eng.json
done
