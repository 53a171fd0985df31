#!/bin/bash
set -u
cd "$(dirname "$0")/.."
CMD="python scripts/extra_bench.py twostage 32"
$CMD > gpurun_out/plain_ts.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_ts.csv $CMD > gpurun_out/ncu_ts.log 2>&1
echo rc=$?; tail -2 gpurun_out/plain_ts.log | cut -c1-200
