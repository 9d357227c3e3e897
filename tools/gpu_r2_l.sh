#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sketch.py -m gpu -q 2>&1 | tail -6 | cut -c1-300
timeout 300 python bench.py --workload sketch_dense --dtype f64 --m 4194304 --d 256 --steps 3 --warmup 3 > gpurun_out/sec_sketch_dense_r2.json 2>gpurun_out/sec_sd.err; python -c "
import json; d=json.load(open('gpurun_out/sec_sketch_dense_r2.json')); print('dense sketch d=256', d['ms_per_step'], d['value']/1e3, d['roofline']['frac'])"; tail -2 gpurun_out/sec_sd.err
timeout 300 python bench.py --workload sketch_dense --dtype f64 --m 4194304 --d 128 --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dense sketch d=128', d['ms_per_step'], d['value']/1e3)"
