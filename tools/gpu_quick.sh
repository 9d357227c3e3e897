#!/bin/bash
# tests + headline bench (no ncu)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for eng in i8 dmma; do
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --engine $eng > gpurun_out/bench_quick_$eng.json 2> gpurun_out/bench_quick.err
tail -2 gpurun_out/bench_quick.err; cut -c1-1500 gpurun_out/bench_quick_$eng.json
done
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --p 0 > gpurun_out/bench_quick_p0.json 2>> gpurun_out/bench_quick.err
cut -c1-1500 gpurun_out/bench_quick_p0.json
