#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python bench.py --workload cqrrpt --engine i8 --steps 2 --warmup 1 > gpurun_out/sec_cqrrpt_i8.json 2> gpurun_out/sec.err; tail -2 gpurun_out/sec.err; cut -c1-400 gpurun_out/sec_cqrrpt_i8.json
python bench.py --workload cqrrpt --engine i8 --dtype f64 --m 4194304 --steps 2 --warmup 1 > gpurun_out/sec_cqrrpt_i8_f64.json 2> gpurun_out/sec.err; tail -2 gpurun_out/sec.err; cut -c1-400 gpurun_out/sec_cqrrpt_i8_f64.json
