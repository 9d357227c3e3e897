#!/bin/bash
mkdir -p gpurun_out
for prio in 1 0; do
RLB200_OZ_AUX_PRIO=$prio RLB200_OZ_ASSUME_CONST=1 timeout 300 python tools/bench_gemm.py 22 1024 256 i8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('prio=$prio', {k:v['ms'] for k,v in d.items() if isinstance(v,dict)})"
done
RLB200_OZ_AUX_PRIO=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rsvd prio=1', d['ms_per_step'], d['value'])"
