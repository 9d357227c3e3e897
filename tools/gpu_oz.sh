#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ozaki.py -m gpu -q -x 2>&1 | tail -3
RLB200_OZ_ASSUME_CONST=1 RLB200_OZ_TIMELINE=1 timeout 600 python tools/bench_gemm.py 22 1024 256 > gpurun_out/bench_gemm_tl.json 2> gpurun_out/bench_gemm_tl.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench_gemm_tl.json")); print({k:v["ms"] for k,v in d.items() if isinstance(v,dict)})
PY
grep "oz timeline] NN" gpurun_out/bench_gemm_tl.err | sed -n 3p | cut -c1-300
grep "oz timeline] TN" gpurun_out/bench_gemm_tl.err | sed -n 3p | cut -c1-300
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rsvd', d['ms_per_step'], d['value'], d['roofline']['class_ms_per_step'])"
