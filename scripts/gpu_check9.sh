#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --pipe-group 4096"
echo "== ncu full fused @4096ch"
$CMD > gpurun_out/plain8.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_block_fused -s 8 -c 2 -o gpurun_out/prof_fused_4096 -f $CMD > gpurun_out/ncu_fused.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/ncu_fused.log
echo "== ncu launch list"
$CMD > gpurun_out/plain9.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_fused_4096.csv $CMD > gpurun_out/ncu_list4.log 2>&1
echo "rc=$?"
