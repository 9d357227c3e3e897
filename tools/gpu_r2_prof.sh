#!/bin/bash
# ncu evidence of the round's final kernels (a number printed by a run under ncu is never a bench value):
#  1. launch list (time + DRAM bytes per launch) of one bench step at 2^20 rows;
#  2. --set full captures of the dominant launches (oz3 A*Omega, oz2 A^T Y) and of the two kernels written this round
#     (qr_coop_kernel: single-launch QRCP of the sketch; saso_strip_kernel: count-sketch).
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/launches_dram_r2b.csv python bench.py --steps 1 --warmup 1 --m 1048576 --no-cpu --no-e2e > gpurun_out/prof_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz3_kernel -s 2 -c 1 -o gpurun_out/prof_oz3_nn_r2i -f \
  python tools/bench_gemm.py 21 1024 256 i8 > gpurun_out/prof_c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz2_kernel -s 4 -c 1 -o gpurun_out/prof_oz2_tn_r2i -f \
  python tools/bench_gemm.py 21 1024 256 i8 > gpurun_out/prof_d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qr_coop_kernel -s 1 -c 1 -o gpurun_out/prof_qr_coop_r2 -f \
  python tools/bench_qrcp.py 4096 2048 f32 > gpurun_out/prof_e.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:saso_strip_kernel -s 2 -c 1 -o gpurun_out/prof_saso_strip_r2 -f \
  python bench.py --workload sketch_sparse --dtype f32 --nnz 1 --steps 2 --warmup 2 > gpurun_out/prof_f.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_dram_r2b.csv | tail -8
