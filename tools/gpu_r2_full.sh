#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2 | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_default_r2.json 2> gpurun_out/bench_default_r2.err
wc -l gpurun_out/bench_default_r2.json
python -c "import json; d=json.load(open('gpurun_out/bench_default_r2.json')); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e'], d['cpu_baseline'], d['parity'], d['clocks'], d['gpu_launches'])" | cut -c1-1500
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm_r2.json 2> gpurun_out/bench_reference_arm_r2.err
head -c 600 gpurun_out/bench_reference_arm_r2.json
