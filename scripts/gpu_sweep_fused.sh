#!/bin/bash
set -u
cd "$(dirname "$0")/.."
for cfg in "4 2" "4 3" "2 2" "2 3" "2 4" "2 6" "8 2" "1 4" "1 8"; do
set -- $cfg
timeout 600 python bench.py --fused-rows $1 --fused-stages $2 --no-cpu-baseline --steps 150 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']
print('rows $1 stages $2 value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'GB/s', round(r['achieved']), 'e2e', round(d['e2e']['value']))" | tee -a gpurun_out/sweep_fused.txt
done
