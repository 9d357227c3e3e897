#!/bin/bash
# One GPU-box visit: tests, the headline bench (both engines), the reference arm, the ncu launch list of the same command and
# full captures of the tensor-core kernel and the slicers.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python bench.py --steps 3 --warmup 3 --p 0 --no-cpu --no-e2e > gpurun_out/bench_1gpu_p0.json 2>> gpurun_out/bench_1gpu.err
python bench.py --steps 3 --warmup 3 --engine dmma --no-cpu --no-e2e > gpurun_out/bench_1gpu_dmma.json 2>> gpurun_out/bench_1gpu.err
python bench.py --steps 3 --warmup 3 --digits 7 --no-cpu --no-e2e > gpurun_out/bench_1gpu_s7.json 2>> gpurun_out/bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --m 2097152 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ozaki_mma_kernel -s 40 -c 1 -o gpurun_out/prof_oz_nn -f \
    python tools/bench_gemm.py 20 1024 256 nn > gpurun_out/ncu_oz_nn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ozaki_mma_kernel -s 10 -c 1 -o gpurun_out/prof_oz_tn -f \
    python tools/bench_gemm.py 20 1024 256 tn > gpurun_out/ncu_oz_tn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:oz_slice -s 40 -c 3 -o gpurun_out/prof_oz_slice -f \
    python tools/bench_gemm.py 20 1024 256 i8 > gpurun_out/ncu_oz_slice.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cut -c1-600 gpurun_out/bench_1gpu.json; cut -c1-300 gpurun_out/bench_ref.json
