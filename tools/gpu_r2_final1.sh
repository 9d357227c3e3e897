#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_final.json')); r=d['roofline']
print('rsvd ms', d['ms_per_step'], 'TF', d['value']/1e3, 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
print('roofline', r['achieved'], r['peak'], r['frac'], r['whole_step_frac'], 'traffic', r['traffic'])
print('class', r['class_ms_per_step']); print('parity', d['parity']); print('clocks', d['clocks'])
PY
tail -2 gpurun_out/bench_r2_final.err
timeout 600 python bench.py --stab plul --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/bench_r2_plul.json 2> gpurun_out/bench_r2_plul.err
python -c "import json; d=json.load(open('gpurun_out/bench_r2_plul.json')); print('PLUL stack', d['ms_per_step'], d['value']/1e3, d['roofline']['class_ms_per_step'])"; tail -2 gpurun_out/bench_r2_plul.err
timeout 600 python bench.py --p 0 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r2_p0.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/bench_r2_p0.json')); print('p=0', d['ms_per_step'], d['value']/1e3)"
