#!/bin/bash
# 2 GPUs: sharded parity (TSQR, both data planes, CQRRT), default bench at N=2, configs[4] shape at N=2 (weak: 2^25 rows per GPU)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_r2_multi.log; tail -5 gpurun_out/pytest_r2_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_2gpu_r2.json 2> gpurun_out/bench_2gpu_r2.err
python -c "import json; d=json.load(open('gpurun_out/bench_2gpu_r2.json')); print('N=2 c2', d['ms_per_step'], d['value']/1e3, 'TF')"; tail -2 gpurun_out/bench_2gpu_r2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --config c5 --steps 2 --warmup 3 > gpurun_out/bench_2gpu_c5_r2.json 2> gpurun_out/bench_2gpu_c5_r2.err
python -c "import json; d=json.load(open('gpurun_out/bench_2gpu_c5_r2.json')); print('N=2 c5', d['ms_per_step'], d['value']/1e3, 'TF', d['scaling'], d['roofline']['class_ms_per_step'])"; tail -2 gpurun_out/bench_2gpu_c5_r2.err
timeout 600 python bench.py --config c5 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_1gpu_c5_r2.json 2> gpurun_out/bench_1gpu_c5_r2.err
python -c "import json; d=json.load(open('gpurun_out/bench_1gpu_c5_r2.json')); print('N=1 c5', d['ms_per_step'], d['value']/1e3, 'TF', d['scaling'], d['roofline']['class_ms_per_step'])"; tail -2 gpurun_out/bench_1gpu_c5_r2.err
