#!/bin/bash
# N-GPU visit: sharded parity test + the scaling bench at N ranks (weak scaling: 2^24 rows per GPU)
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_multi.log
tail -4 gpurun_out/pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 3 --warmup 3 \
  > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -3 gpurun_out/bench_${N}gpu.err; cut -c1-700 gpurun_out/bench_${N}gpu.json
