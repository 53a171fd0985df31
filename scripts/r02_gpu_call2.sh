#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest2.log
tail -40 gpurun_out/r02_pytest2.log
