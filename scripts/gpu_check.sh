#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err
python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); r=d['roofline']
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'GB/s', round(r['achieved']), 'frac', round(r['frac'],3), 'traffic', r['traffic'], 'e2e', round(d['e2e']['value']), 'e2e_ms', round(d['e2e']['ms_per_step'],4), 'cpu', round(d['cpu_baseline']['value']), d['cpu_baseline']['cores'])"
echo "== configs"; timeout 800 python scripts/configs_bench.py 2>&1 | tee gpurun_out/configs.jsonl | cut -c1-330
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 3 2>/dev/null | tail -1 | cut -c1-300
echo "== ncu launch list (default bench)"
python bench.py --steps 6 --warmup 3 > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 6 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python scripts/ncu_kernel_times.py gpurun_out/launches_bench.csv | head -8
