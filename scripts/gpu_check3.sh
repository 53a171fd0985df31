#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
echo "== bench (default)"; timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
echo "== pipe group sweep"
rm -f gpurun_out/sweep_pipe.txt
for g in 128 256 512 1024; do
  timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --pipe-group $g 2>/dev/null \
    | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pipe_group $g', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e_ms', round(d['e2e']['ms_per_step'],4))" | tee -a gpurun_out/sweep_pipe.txt
done
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --pipe-group 4096"
echo "== ncu full K2 @4096ch"
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_mac -s 8 -c 2 -o gpurun_out/prof_k2_4096 -f $CMD > gpurun_out/ncu_full2.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_full2.log
