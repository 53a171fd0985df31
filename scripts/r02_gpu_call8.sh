#!/bin/bash
set -x
mkdir -p gpurun_out
/usr/bin/gcc -O2 -Iinclude tests/abi/latency.c -o /tmp/latency -Lfft_convolution_b200 -lfftconv_b200 -Wl,-rpath,$PWD/fft_convolution_b200 -lm
/tmp/latency 1 300 > /dev/null && ncu --set full --clock-control none --import-source on -k regex:k_block_fused_pair -s 300 -c 1 -f -o gpurun_out/r02_pair_cfg1 /tmp/latency 1 300 > /dev/null 2>&1
/tmp/latency 0 300 > /dev/null && ncu --set full --clock-control none --import-source on -k regex:k_block_fused -s 300 -c 1 -f -o gpurun_out/r02_fused_cfg0 /tmp/latency 0 300 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
