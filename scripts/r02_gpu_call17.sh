#!/bin/bash
set -x
mkdir -p gpurun_out
for b in 4096 8192 16384; do python scripts/bigfft_probe.py $b > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 25 -c 45 --csv --log-file gpurun_out/r02_bigfft_$b.csv python scripts/bigfft_probe.py $b > /dev/null 2>&1; done
