#!/bin/bash
# secondary workloads (BASELINE configs[2], configs[3], sketch micro-kernels) through bench.py: full lines with clocks, roofline, cpu_baseline
mkdir -p gpurun_out
timeout 900 python bench.py --workload cqrrpt --steps 2 --warmup 3 > gpurun_out/sec_cqrrpt_c3_r2.json 2> gpurun_out/sec_c3.err; python -c "
import json; d=json.load(open('gpurun_out/sec_cqrrpt_c3_r2.json')); print('C3 cqrrpt', d['ms_per_step'], d['value']/1e3, d['roofline']['frac'], d.get('class_ms_per_step'), d['cpu_baseline'])"
tail -2 gpurun_out/sec_c3.err
timeout 900 python bench.py --workload cqrrpt --dtype f64 --m 4194304 --steps 2 --warmup 2 --no-cpu > gpurun_out/sec_cqrrpt_f64_r2.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/sec_cqrrpt_f64_r2.json')); print('cqrrpt f64 2^22', d['ms_per_step'], d['value']/1e3, d['roofline']['frac'], d.get('class_ms_per_step'))"
timeout 1200 python bench.py --workload bqrrp --n 32768 --steps 1 --warmup 1 --no-cpu > gpurun_out/sec_bqrrp_32k_r2.json 2> gpurun_out/sec_bq.err; python -c "
import json; d=json.load(open('gpurun_out/sec_bqrrp_32k_r2.json')); print('bqrrp 32k', d['ms_per_step'], d['value']/1e3, d['roofline']['frac'], d.get('class_ms_per_step'))"
tail -2 gpurun_out/sec_bq.err
timeout 1500 python bench.py --workload bqrrp --steps 1 --warmup 1 > gpurun_out/sec_bqrrp_c4_r2.json 2> gpurun_out/sec_bq4.err; python -c "
import json; d=json.load(open('gpurun_out/sec_bqrrp_c4_r2.json')); print('C4 bqrrp 65536', d['ms_per_step'], d['value']/1e3, d['roofline']['frac'], d.get('class_ms_per_step'), d['cpu_baseline'])"
tail -2 gpurun_out/sec_bq4.err
for nnz in 1 2 4; do timeout 300 python bench.py --workload sketch_sparse --nnz $nnz --steps 5 --warmup 3 > gpurun_out/sec_sketch_sparse_nnz${nnz}_r2.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/sec_sketch_sparse_nnz${nnz}_r2.json')); print('sparse nnz$nnz', d['ms_per_step'], d['value'], d['roofline']['frac'])"; done
timeout 300 python bench.py --workload sketch_dense --dtype f64 --m 4194304 --d 256 --steps 3 --warmup 3 > gpurun_out/sec_sketch_dense_r2.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/sec_sketch_dense_r2.json')); print('dense sketch', d['ms_per_step'], d['value']/1e3, d['roofline']['frac'])"
