#!/bin/bash
for f in 0 128 4 6; do echo "== oz3 dbg=$f"; RLB200_OZ2_DBG=$f RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 nn 2>&1 | cut -c1-200 | tail -1; done
