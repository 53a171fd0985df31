#!/usr/bin/env python
"""K1 / K2 / K3 at a large block size (the two-stage tail): 4096 channels, block 8192, 10 segments; run under
ncu --metrics gpu__time_duration.sum for the per-kernel times."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fft_convolution_b200 as F

import os
from fft_convolution_b200 import _lib
for kv in filter(None, os.environ.get("FCB_TUNE", "").split(",")):
    k, v = kv.split("=")
    _lib.check(_lib.load().fcb_tune(k.encode(), int(v)))
C, B, S = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 10
rng = np.random.default_rng(0)
h = (rng.standard_normal((C, B * S)) * 0.01).astype(np.float32)
st = torch.cuda.Stream()
g = F.FFTConvolver.init(h, B, B * S, stream=st.cuda_stream)
x = torch.rand((C, B), device="cuda") - 0.5
y = torch.empty((C, B), device="cuda")
for _ in range(6):
    g.process_dev(x.data_ptr(), B, B, y.data_ptr(), B, B)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(10):
    g.process_dev(x.data_ptr(), B, B, y.data_ptr(), B, B)
e1.record(st)
torch.cuda.synchronize()
print(f"B={B}: {e0.elapsed_time(e1) / 10:.3f} ms per block (K1+K2+K3), 4096 channels")
