#!/bin/bash
# GPU tests, headline bench with the pipelined e2e path, pipe-group sweep, ncu K2 capture at the
# bench's own channel count.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
echo "== bench (default)"; timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
echo "== pipe group sweep"
for g in 128 256 512 1024 2048 4096; do
  timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --pipe-group $g 2>/dev/null \
    | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pipe_group $g', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e_ms', round(d['e2e']['ms_per_step'],4))" | tee -a gpurun_out/sweep_pipe.txt
done
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
echo "== ncu full K2 @4096ch"
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_mac -s 12 -c 2 -o gpurun_out/prof_k2_4096 -f $CMD > gpurun_out/ncu_full2.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_full2.log
echo "== ncu launch list @4096ch"
$CMD > gpurun_out/plain4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_4096.csv $CMD > gpurun_out/ncu_list2.log 2>&1
echo "rc=$?"
