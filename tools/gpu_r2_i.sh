#!/bin/bash
# full GPU test-suite, smoke, the default bench line, and the DRAM-bytes launch list of one step at m = 2^20
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_r2i_all.log; tail -6 gpurun_out/pytest_r2i_all.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2i.json'))
r=d['roofline']
print('rsvd ms', d['ms_per_step'], 'TF', d['value']/1e3, 'e2e', d['e2e'], 'cpu', d['cpu_baseline'])
print('roofline achieved', r['achieved'], 'peak', r['peak'], 'frac', r['frac'], 'whole', r['whole_step_frac'], r['peak_source'])
print('class', r['class_ms_per_step'])
print('parity', d['parity'])
print('clocks', d['clocks'])
PY
tail -3 gpurun_out/bench_r2i.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_dram_r2.csv python bench.py --m 1048576 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_ncu_r2i.log 2>&1
tail -2 gpurun_out/bench_ncu_r2i.log | cut -c1-300; wc -l gpurun_out/launches_dram_r2.csv
