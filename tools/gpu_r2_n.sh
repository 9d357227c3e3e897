#!/bin/bash
# timing experiments on the TN launch: which feed limits it (RLB200_OZ2_DBG: 2 = no loads of A, 64 = no bulk copies of the Y digits, 8 = no conversion math)
for f in 0 64 2 66 74; do
echo "== dbg $f"; RLB200_OZ2_DBG=$f RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | cut -c1-250 | tail -1
done
echo "== NN on oz2 (non persistent)"; 
for f in 0 64 2 66; do
echo "== dbg $f"; RLB200_OZ3=0 RLB200_OZ2_DBG=$f RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | cut -c1-250 | tail -1
done
