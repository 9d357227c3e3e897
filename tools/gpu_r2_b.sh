#!/bin/bash
mkdir -p gpurun_out
timeout 300 tools/peaks_i8 1 > gpurun_out/peaks_i8_r2b.json 2> gpurun_out/peaks_i8_r2b.err; echo "peaks rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/peaks_i8_r2b.json'))
for r in d['results']:
    print(r['cta_group'], r['pattern'], 'instr', r['instr_per_step'], 'cyc/step', r['cycles_per_step'], 'floor', r['floor_cycles_per_step'], 'burst', r['burst_tops'], 'sus', r['sustained_tops'], r['sustained_ghz'])
PY
tail -3 gpurun_out/peaks_i8_r2b.err
timeout 600 python tools/dbg_headline.py 2>&1 | tail -30 | tee gpurun_out/dbg_headline_r2b.log
