#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
/usr/bin/gcc -O2 -Iinclude tests/abi/latency.c -o /tmp/latency -Lfft_convolution_b200 -lfftconv_b200 -Wl,-rpath,$PWD/fft_convolution_b200 -lm
for c in 0 1 2; do /tmp/latency $c 3000; /tmp/latency $c 3000 mapped_io=0; done 2>&1 | tee gpurun_out/r02_abi_latency2.jsonl
python scripts/r02_small_batch_probe.py 2>&1 | grep shape
