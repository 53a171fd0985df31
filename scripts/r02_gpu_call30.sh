#!/bin/bash
set -x
/usr/bin/gcc -O2 -Iinclude tests/abi/latency.c -o /tmp/latency -Lfft_convolution_b200 -lfftconv_b200 -Wl,-rpath,$PWD/fft_convolution_b200 -lm
for st in 2 1 3; do for c in 0 1; do /tmp/latency $c 3000 split_min_stages=$st | cut -c1-140; done; done
