#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --config c5 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_8gpu_c5_r2.json 2> gpurun_out/bench_8gpu_c5_r2.err
wc -l gpurun_out/bench_8gpu_c5_r2.json; python -c "import json; d=json.load(open('gpurun_out/bench_8gpu_c5_r2.json')); print('N=8 c5', d['scaling'], d['ms_per_step'], d['value']/1e3, 'TF', d['roofline']['class_ms_per_step'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_8gpu_r2.json 2> gpurun_out/bench_8gpu_r2.err
wc -l gpurun_out/bench_8gpu_r2.json; python -c "import json; d=json.load(open('gpurun_out/bench_8gpu_r2.json')); print('N=8 c2 weak', d['ms_per_step'], d['value']/1e3, 'TF')"
