#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python scripts/r02_update_bench.py 2>&1 | tee gpurun_out/r02_update_bench.jsonl
