#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
for f in 1 0; do
echo "== bench fused=$f"; timeout 900 python bench.py --fused $f --no-cpu-baseline > gpurun_out/bench_fused$f.json 2> gpurun_out/bench_fused$f.err; tail -2 gpurun_out/bench_fused$f.err
python -c "
import json; d=json.load(open('gpurun_out/bench_fused$f.json')); r=d['roofline']
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(r['avg_launch_ms'],4), 'GB/s', round(r['achieved']), 'frac', round(r['frac'],3), 'share', round(r['kernel_share_of_step'],3), 'e2e', round(d['e2e']['value']), 'e2e_ms', round(d['e2e']['ms_per_step'],4), 'launches', d['gpu_launches'])"
done
