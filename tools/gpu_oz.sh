#!/bin/bash
mkdir -p gpurun_out
RLB200_OZ_ASSUME_CONST=1 timeout 300 python tools/bench_gemm.py 22 1024 256 i8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:v['ms'] for k,v in d.items() if isinstance(v,dict)})"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rsvd', d['ms_per_step'], d['value'])"
