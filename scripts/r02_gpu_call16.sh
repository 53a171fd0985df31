#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_realtime.py -m gpu -q -x 2>&1 | tail -5
for b in 4096 8192 16384; do python scripts/bigfft_probe.py $b > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 18 --csv --log-file gpurun_out/r02_bigfft_$b.csv python scripts/bigfft_probe.py $b > /dev/null 2>&1; done
python scripts/extra_bench.py twostage 320 2>&1 | tail -1
