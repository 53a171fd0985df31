#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for st in 2 3; do for pg in 256 512 1024; do
timeout 900 python bench.py --fused-stages $st --pipe-group $pg --no-cpu-baseline --steps 200 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']
print('stages $st pipe $pg value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'GB/s', round(r['achieved']), 'frac', round(r['frac'],3), 'e2e', round(d['e2e']['value']), 'e2e_ms', round(d['e2e']['ms_per_step'],4))"
done; done
