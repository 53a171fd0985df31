#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
/usr/bin/gcc -O2 -Iinclude tests/abi/latency.c -o /tmp/latency -Lfft_convolution_b200 -lfftconv_b200 -Wl,-rpath,$PWD/fft_convolution_b200 -lm
for c in 0 1 2; do /tmp/latency $c 3000; /tmp/latency $c 3000 pinned; /tmp/latency $c 3000 split=0 mapped_io=0; done 2>&1 | tee gpurun_out/r02_abi_latency.jsonl
python bench.py --channels 512 --steps 300 --warmup 20 --realtime 0 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('512ch', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])"
