#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
rm -f gpurun_out/mimo_1gpu.json
for ns in 1 4 16; do
  timeout 600 python scripts/mimo_bench.py --streams $ns 2>>gpurun_out/mimo1.err | tee -a gpurun_out/mimo_1gpu.json
done
