#!/bin/bash
# Switch-off series behind DESIGN.md 3b ("what bounds the kernel"): the A^T Y launch at m = 2^21, n = 1024, k = 256 without cluster sharing.
# RLB200_OZ2_DBG bits: 1 cycle stamps, 2 no loads of A, 8 no conversion arithmetic, 32 no MMAs, 64 no bulk copies of the Y digits,
# 256 no digit stores, 512 no proxy fence in the issuer, 1024 wide instructions last.  Run on a GPU box: bash tools/gpu_r2_n.sh
for f in 0 2 66 322 330 331 363 1354; do
echo "== RLB200_OZ2_SHARE=0 RLB200_OZ2_DBG=$f"
RLB200_OZ2_SHARE=0 RLB200_OZ2_DBG=$f RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | grep -a "i8s6\|oz2 dbg" | tail -2 | cut -c1-420
done
for ng in 1 2 3; do
echo "== converter groups $ng, everything off"
RLB200_OZ2_NG=$ng RLB200_OZ2_SHARE=0 RLB200_OZ2_DBG=330 RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | grep -a "i8s6" | tail -1 | cut -c100-300
done
echo "== persistent A*Omega kernel: two stages per issue round (default) vs one"
RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | grep -a "i8s6" | tail -1 | cut -c1-300
RLB200_OZ3_ONE_STAGE=1 RLB200_OZ_ASSUME_CONST=1 timeout 120 python tools/bench_gemm.py 21 1024 256 i8 2>&1 | grep -a "i8s6" | tail -1 | cut -c1-300
