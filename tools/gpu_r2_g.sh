#!/bin/bash
mkdir -p gpurun_out
for dbg in 16 0; do
for args in "640 256 256 6" "640 256 256 7" "640 1024 256 7" "5000 1024 256 7" "640 256 512 6"; do
  echo "== DBG=$dbg $args"; RLB200_OZ2_DBG=$dbg CUDA_LAUNCH_BLOCKING=1 timeout 40 python tools/dbg_share.py $args 2>&1 | tail -4 | cut -c1-300; echo "rc=$?"
done; done
