#!/bin/bash
set -x
mkdir -p gpurun_out
python scripts/bigfft_probe.py 8192 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_rfft_forward|k_irfft_ola' -s 24 -c 2 -f -o gpurun_out/r02_tailfft_wide python scripts/bigfft_probe.py 8192 > /dev/null 2>&1
