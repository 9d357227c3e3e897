#!/bin/bash
# 2 GPUs: sharded parity (TSQR, both data planes, CQRRT, CQRRPT, RSVD at k >= 64), default bench at N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_r2_multi.log; tail -3 gpurun_out/pytest_r2_multi.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu_r2.json 2> gpurun_out/bench_2gpu_r2.err
wc -l gpurun_out/bench_2gpu_r2.json; python -c "import json; d=json.load(open('gpurun_out/bench_2gpu_r2.json')); print('N=2 c2', d['ms_per_step'], d['value']/1e3, 'TF')"
