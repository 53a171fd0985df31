#!/bin/bash
# round 2, GPU call 1: test suite + ncu --set full of the 16384-point tail transforms and the paired kernel
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
tail -5 gpurun_out/r02_pytest1.log
python scripts/bigfft_probe.py 8192 > gpurun_out/r02_bigfft_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_rfft_forward|k_irfft_ola' -s 8 -c 4 -f -o gpurun_out/r02_tailfft python scripts/bigfft_probe.py 8192 > gpurun_out/r02_tailfft_ncu.log 2>&1
cat gpurun_out/r02_bigfft_plain.log
python scripts/extra_bench.py twostage 64 > gpurun_out/r02_ts_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_block_fused_pair -s 40 -c 2 -f -o gpurun_out/r02_pair python scripts/extra_bench.py twostage 64 > gpurun_out/r02_pair_ncu.log 2>&1
cat gpurun_out/r02_ts_plain.log
python scripts/configs_bench.py > gpurun_out/r02_configs_before.jsonl 2>&1; cat gpurun_out/r02_configs_before.jsonl
