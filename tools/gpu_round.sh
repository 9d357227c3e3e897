#!/bin/bash
# One GPU-box visit: tests, the headline bench, the ncu launch list of the same command and full captures of the two tall GEMMs.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_nn_kernel -s 2 -c 2 -o gpurun_out/prof_nn -f \
    python bench.py --m 2097152 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_nn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tn_kernel -s 2 -c 2 -o gpurun_out/prof_tn -f \
    python bench.py --m 2097152 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_tn.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_1gpu.json; cat gpurun_out/bench_ref.json
