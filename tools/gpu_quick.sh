#!/bin/bash
# tests + headline bench (no ncu)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --p 0 > gpurun_out/bench_quick_p0.json 2>> gpurun_out/bench_quick.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_quick.json gpurun_out/bench_quick_p0.json | cut -c1-2200
