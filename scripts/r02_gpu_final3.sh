#!/bin/bash
# last capture of round 2: suite, smoke, bench line, reference arm, launch list, hash-stamped DRAM traffic
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_gpu_suite.txt 2>&1; tail -4 gpurun_out/r02_gpu_suite.txt
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/r02_smoke.txt
( time timeout 400 python bench.py ) > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -4 gpurun_out/r02_bench_1gpu.err
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -4 gpurun_out/r02_bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 5 --warmup 3 --realtime 0 --no-cpu-baseline > /dev/null 2>&1
timeout 400 bash scripts/capture_traffic.sh 2>&1 | tail -2
