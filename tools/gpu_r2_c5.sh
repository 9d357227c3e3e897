#!/bin/bash
# BASELINE configs[4]: 128M x 512, k = 128, strong scaling at N = 4 (2^25 rows per GPU)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --config c5 --steps 3 --warmup 3 > gpurun_out/bench_4gpu_c5_r2.json 2> gpurun_out/bench_4gpu_c5_r2.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_4gpu_c5_r2.json') if l.startswith('{')][0]); print('N=4 c5', d['ms_per_step'], d['value']/1e3, 'TF', d['scaling'], d['config']['m_per_gpu'], d['roofline']['class_ms_per_step'])"
tail -2 gpurun_out/bench_4gpu_c5_r2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_4gpu_r2.json 2> gpurun_out/bench_4gpu_r2.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_4gpu_r2.json') if l.startswith('{')][0]); print('N=4 c2 weak', d['ms_per_step'], d['value']/1e3, 'TF')"
