#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err
python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); r=d['roofline']
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'k2_ms', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],3), 'share', round(r['kernel_share_of_step'],3), 'e2e', round(d['e2e']['value']), 'cpu', d['cpu_baseline']['value'] if d['cpu_baseline'] else None)"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --pipe-group 4096"
$CMD > gpurun_out/plain7.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_4096.csv $CMD > gpurun_out/ncu_list3.log 2>&1
echo "rc=$?"
