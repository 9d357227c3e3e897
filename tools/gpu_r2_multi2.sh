#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_2gpu_r2.json 2> gpurun_out/bench_2gpu_r2.err
wc -l gpurun_out/bench_2gpu_r2.json; head -c 300 gpurun_out/bench_2gpu_r2.json; echo
python -c "import json; d=json.load(open('gpurun_out/bench_2gpu_r2.json')); print('N=2 c2', d['ms_per_step'], d['value']/1e3, 'TF', d['e2e'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_2gpu_ref_r2.json 2> gpurun_out/bench_2gpu_ref_r2.err
wc -l gpurun_out/bench_2gpu_ref_r2.json; head -c 400 gpurun_out/bench_2gpu_ref_r2.json; echo
