#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
/usr/bin/gcc -O2 -Iinclude tests/abi/latency.c -o /tmp/latency -Lfft_convolution_b200 -lfftconv_b200 -Wl,-rpath,$PWD/fft_convolution_b200 -lm
for c in 0 1; do /tmp/latency $c 3000; /tmp/latency $c 3000 k1_late=0; done 2>&1
for k in 1 0; do python - $k <<'PY'
import sys, json, subprocess
k = sys.argv[1]
sys.path.insert(0, '.')
from fft_convolution_b200 import _lib
_lib.check(_lib.load().fcb_tune(b"k1_late", int(k)))
import bench
sys.argv = ['bench.py', '--steps', '300', '--warmup', '20', '--realtime', '0', '--no-cpu-baseline']
import io, contextlib
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    bench.main()
d = json.loads(buf.getvalue().strip().splitlines()[-1])
print('k1_late', k, 'device ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'])
PY
done
