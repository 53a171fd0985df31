#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_realtime.py -m gpu -q -x 2>&1 | tail -4
/usr/bin/gcc -O2 -Iinclude tests/abi/latency.c -o /tmp/latency -Lfft_convolution_b200 -lfftconv_b200 -Wl,-rpath,$PWD/fft_convolution_b200 -lm
for c in 0 1 2; do /tmp/latency $c 3000; /tmp/latency $c 3000 pinned; /tmp/latency $c 3000 split=0 mapped_io=0; done 2>&1 | tee gpurun_out/r02_abi_latency.jsonl
python scripts/configs_bench.py > gpurun_out/r02_configs.jsonl 2>&1; cat gpurun_out/r02_configs.jsonl
