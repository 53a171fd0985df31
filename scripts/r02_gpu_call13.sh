#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_8gpu.txt 2>&1; numactl -H >> gpurun_out/r02_topo_8gpu.txt 2>&1; lscpu | head -25 >> gpurun_out/r02_topo_8gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
( time $TR bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; tail -c 1500 gpurun_out/r02_bench_8gpu.json; tail -5 gpurun_out/r02_bench_8gpu.err
( time $TR bench.py --gpus 8 --steps 200 --warmup 10 --zero-copy 0 --realtime 0 --mimo 0 ) > gpurun_out/r02_bench_8gpu_copy_engines.json 2> gpurun_out/r02_bench_8gpu_copy_engines.err; tail -c 600 gpurun_out/r02_bench_8gpu_copy_engines.json
( time $TR bench.py --gpus 8 --steps 200 --warmup 10 --realtime 0 --mimo 0 ) > gpurun_out/r02_bench_8gpu_zero_copy.json 2> gpurun_out/r02_bench_8gpu_zero_copy.err; tail -c 600 gpurun_out/r02_bench_8gpu_zero_copy.json
