#!/bin/bash
set -x
mkdir -p gpurun_out
/usr/bin/gcc -O2 -Iinclude tests/abi/latency.c -o gpurun_out/latency -Lfft_convolution_b200 -lfftconv_b200 -Wl,-rpath,$PWD/fft_convolution_b200 -lm
./gpurun_out/latency 2 3000
for c in 0 1 2; do
./gpurun_out/latency $c 300 > /dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 24 --csv --log-file gpurun_out/r02_lat_cfg$c.csv ./gpurun_out/latency $c 300 > /dev/null 2>&1
done
python -m pytest tests/test_gpu_realtime.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
