#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest4.log 2>&1; tail -15 gpurun_out/r02_pytest4.log
python scripts/configs_bench.py > gpurun_out/r02_configs_split.jsonl 2> gpurun_out/r02_configs_split.err; cat gpurun_out/r02_configs_split.jsonl; tail -3 gpurun_out/r02_configs_split.err
FCB_TUNE_SPLIT=0 python - <<'PY' > gpurun_out/r02_configs_nosplit.jsonl 2>&1
import sys, json
sys.path.insert(0, 'scripts'); sys.path.insert(0, '.')
from fft_convolution_b200 import _lib
_lib.check(_lib.load().fcb_tune(b"split", 0))
import configs_bench as cb
for fn in (cb.cfg0, lambda: cb.cfg1(True), cb.cfg2):
    print(json.dumps(fn()), flush=True)
PY
cat gpurun_out/r02_configs_nosplit.jsonl
