#!/bin/bash
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
for k in 1 0 1 0; do
$TR bench.py --gpus 8 --steps 300 --warmup 20 --realtime 0 --mimo 0 --k1-late $k 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'k1_late': $k, 'n_gpus': d['n_gpus'], 'device_ms': d['ms_per_step'], 'e2e_ms': d['e2e']['ms_per_step'], 'e2e_value': d['e2e']['value'], 'value': d['value']}))" | tee -a gpurun_out/r02_e2e_8gpu_k1_late.jsonl
done
