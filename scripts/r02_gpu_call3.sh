#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_realtime.py tests/test_abi_replay.py -m gpu -q -x > gpurun_out/r02_pytest3.log 2>&1; tail -3 gpurun_out/r02_pytest3.log
( time python bench.py ) > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -c 6000 gpurun_out/r02_bench_1gpu.json; tail -5 gpurun_out/r02_bench_1gpu.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -c 1500 gpurun_out/r02_bench_ref.json
bash scripts/capture_traffic.sh 2>&1 | tail -3
