#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/sec_$name.json 2> gpurun_out/sec_$name.err; echo "== $name rc=$?"; tail -c 1500 gpurun_out/sec_$name.json; tail -3 gpurun_out/sec_$name.err; }
run sketch_sparse_nnz1 --workload sketch_sparse --nnz 1 --steps 5 --warmup 3
run sketch_sparse_nnz4 --workload sketch_sparse --nnz 4 --steps 5 --warmup 3
run sketch_sparse_f64 --workload sketch_sparse --nnz 2 --dtype f64 --m 4194304 --steps 5 --warmup 3
run sketch_dense --workload sketch_dense --dtype f64 --d 256 --m 4194304 --n 2048 --steps 3 --warmup 3
run cqrrpt_c3 --workload cqrrpt --steps 2 --warmup 1
run bqrrp_16k --workload bqrrp --n 16384 --steps 1 --warmup 1
run bqrrp_32k --workload bqrrp --n 32768 --steps 1 --warmup 0
