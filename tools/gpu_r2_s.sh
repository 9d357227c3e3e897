#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_cqrrpt.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-300
for nnz in 1 2 4; do
timeout 300 python bench.py --workload sketch_sparse --dtype f32 --nnz $nnz --steps 5 --warmup 3 > gpurun_out/sec_sketch_sparse_nnz${nnz}_r2b.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/sec_sketch_sparse_nnz${nnz}_r2b.json')); print('nnz', $nnz, d['ms_per_step'], d['roofline']['frac'])"
done
timeout 300 python bench.py --workload sketch_sparse --dtype f64 --nnz 1 --steps 5 --warmup 3 > gpurun_out/sec_sketch_sparse_f64_r2b.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/sec_sketch_sparse_f64_r2b.json')); print('f64 nnz 1', d['ms_per_step'], d['roofline']['frac'])"
