#!/bin/bash
set -x
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 10 --realtime 0 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'n_gpus': d['n_gpus'], 'device_ms': d['ms_per_step'], 'e2e': d['e2e']}))"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514"
$TR bench.py --gpus 8 --steps 200 --warmup 10 --realtime 0 --mimo 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'n_gpus': d['n_gpus'], 'device_ms': d['ms_per_step'], 'e2e': d['e2e']}))"
