#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_drivers.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for p in 2 0; do
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --p $p > gpurun_out/bench_quick_i8_p$p.json 2> gpurun_out/bench_quick.err
tail -2 gpurun_out/bench_quick.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_quick_i8_p$p.json"))
print("p=$p", d["ms_per_step"], d["value"], d["roofline"]["class_ms_per_step"], d["roofline"]["frac"])
PY
done
RLB200_NO_GRAM_FUSION=1 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('nofusion', d['ms_per_step'], d['value'])"
