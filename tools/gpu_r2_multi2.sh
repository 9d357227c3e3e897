#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/_multi_worker.py 2>/dev/null | grep MULTI_RESULT | python -c "
import json,sys
d=json.loads(sys.stdin.read()[len('MULTI_RESULT '):])
for k,v in d[0].items(): print(k, v)
"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --m 1048576 --steps 2 --warmup 3 2>/dev/null | head -c 300
