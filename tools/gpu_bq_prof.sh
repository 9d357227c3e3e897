#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bqrrp.csv \
    python bench.py --workload bqrrp --n 16384 --steps 1 --warmup 0 > gpurun_out/ncu_bq.log 2>&1
tail -2 gpurun_out/ncu_bq.log | cut -c1-300
