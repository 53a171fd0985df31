#!/bin/bash
# Runs on the GPU box under gpurun: GPU tests, headline bench, K2 sweeps, ncu launch list + one
# full capture of K2.  Everything lands in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.csv
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
echo "== bench (default)"; timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
echo "== sweeps"
for cfg in "1 3" "2 2" "2 3" "2 4" "2 6"; do
  set -- $cfg
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --mac-impl $1 --mac-stages $2 2>/dev/null \
    | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('impl $1 stages $2', 'value', round(d['value']), 'k2 GB/s', round(r['achieved']), 'frac', round(r['frac'],3), 'k2_ms', round(r['avg_launch_ms'],4), 'step_ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))" \
    | tee -a gpurun_out/sweep.txt
done
CMD="python bench.py --steps 2 --warmup 1 --channels 1024 --no-cpu-baseline"
echo "== ncu launch list"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_list.log
echo "== ncu full K2"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_mac -s 8 -c 2 -o gpurun_out/prof_k2 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
