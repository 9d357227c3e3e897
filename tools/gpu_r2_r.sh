#!/bin/bash
# secondary workloads with the final code: BASELINE configs[2] (CQRRPT 2^23 x 2048 fp32), an fp64 CQRRPT case, configs[3] (BQRRP 65536^2) and 32768^2
mkdir -p gpurun_out
timeout 600 python bench.py --workload cqrrpt --dtype f32 --steps 2 --warmup 3 --no-cpu > gpurun_out/sec_cqrrpt_c3_r2c.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/sec_cqrrpt_c3_r2c.json')); print('cqrrpt c3', d['ms_per_step'], d.get('class_ms_per_step'))"
timeout 600 python bench.py --workload cqrrpt --dtype f64 --m 4194304 --steps 2 --warmup 2 --no-cpu > gpurun_out/sec_cqrrpt_f64_r2c.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/sec_cqrrpt_f64_r2c.json')); print('cqrrpt f64', d['ms_per_step'], d['value'])"
timeout 900 python bench.py --workload bqrrp --dtype f64 --n 65536 --steps 1 --warmup 1 --no-cpu > gpurun_out/sec_bqrrp_c4_r2c.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/sec_bqrrp_c4_r2c.json')); print('bqrrp c4', d['ms_per_step'], d['value'], d.get('class_ms_per_step'))"
