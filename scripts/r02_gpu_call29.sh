#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
for b in 4096 8192; do python scripts/bigfft_probe.py $b > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 25 -c 45 --csv --log-file gpurun_out/r02_bigfft_$b.csv python scripts/bigfft_probe.py $b > /dev/null 2>&1; done
python scripts/extra_bench.py twostage 320 2>&1 | tail -1
