#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --pipe-group 4096"
$CMD > gpurun_out/plain5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_irfft_ola -s 8 -c 1 -o gpurun_out/prof_k3 -f $CMD > gpurun_out/ncu_k3.log 2>&1
echo "rc=$?"
$CMD > gpurun_out/plain6.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_rfft_forward -s 16 -c 1 -o gpurun_out/prof_k1 -f $CMD > gpurun_out/ncu_k1.log 2>&1
echo "rc=$?"
