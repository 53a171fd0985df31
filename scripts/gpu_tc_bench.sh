#!/bin/bash
# K4 vs the CUDA-core matrix kernel at the configs[4] shape with many streams
for cfg in "$@"; do
    set -- $cfg
    timeout 600 python scripts/mimo_bench.py --streams $1 --tc $2 --steps $3 --warmup 3 2>>gpurun_out/tc_bench.err | tail -1
done | tee -a gpurun_out/tc_bench.jsonl
