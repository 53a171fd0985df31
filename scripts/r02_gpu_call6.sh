#!/bin/bash
set -x
mkdir -p gpurun_out
/usr/bin/gcc -O2 -Iinclude tests/abi/latency.c -o gpurun_out/latency -Lfft_convolution_b200 -lfftconv_b200 -Wl,-rpath,$PWD/fft_convolution_b200 -lm
for c in 0 1 2; do ./gpurun_out/latency $c 3000; ./gpurun_out/latency $c 3000 split=0; done > gpurun_out/r02_abi_latency.jsonl 2>&1
cat gpurun_out/r02_abi_latency.jsonl
./gpurun_out/latency 0 300 > /dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 12 --csv --log-file gpurun_out/r02_lat_cfg0.csv ./gpurun_out/latency 0 300 > /dev/null 2>&1
./gpurun_out/latency 1 300 > /dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 12 --csv --log-file gpurun_out/r02_lat_cfg1.csv ./gpurun_out/latency 1 300 > /dev/null 2>&1
./gpurun_out/latency 2 300 > /dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 12 --csv --log-file gpurun_out/r02_lat_cfg2.csv ./gpurun_out/latency 2 300 > /dev/null 2>&1
grep -h "k_\|memcpy\|Memcpy" gpurun_out/r02_lat_cfg*.csv | cut -d, -f5,12- | head -40
