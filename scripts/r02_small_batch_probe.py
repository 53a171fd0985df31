#!/usr/bin/env python
"""Where a small-batch block goes: for the configs[0..2] shapes, (a) blocks queued back to back on device buffers (the
kernels' own duration, launch overhead hidden), (b) one block at a time with a host sync after each (launch + kernel +
completion), (c) the host-pointer call (adds the copies in and out).  One JSON line per shape."""
import json, sys, time
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
import fft_convolution_b200 as F

def probe(name, make, C, B):
    st = torch.cuda.Stream()
    conv = make(st.cuda_stream)
    x = torch.from_numpy(bench.synth_noise(0, C, 0, B)).cuda()
    y = torch.empty((C, B), device="cuda")
    xh, yh = bench.synth_noise(0, C, 0, B), np.zeros((C, B), np.float32)
    if C == 1:
        xh, yh = xh[0], yh[0]
    for _ in range(300):
        conv.process_dev(x.data_ptr(), B, B, y.data_ptr(), B, B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 2000
    e0.record(st)
    for _ in range(n):
        conv.process_dev(x.data_ptr(), B, B, y.data_ptr(), B, B)
    e1.record(st)
    torch.cuda.synchronize()
    queued = e0.elapsed_time(e1) / n * 1e3
    t0 = time.perf_counter()
    for _ in range(n):
        conv.process_dev(x.data_ptr(), B, B, y.data_ptr(), B, B)
        conv.sync()
    synced = (time.perf_counter() - t0) / n * 1e6
    t0 = time.perf_counter()
    for _ in range(n):
        conv.process(xh, yh)
    host = (time.perf_counter() - t0) / n * 1e6
    print(json.dumps({"shape": name, "us_per_block_queued_back_to_back": queued, "us_per_block_launch_and_sync": synced,
                      "us_per_block_host_pointer_call_python": host}), flush=True)

h0 = bench.synth_irs(0, 1, 0, 48000)[0]
probe("configs[0] mono B=256 L=48000", lambda s: F.FFTConvolver.init(h0, 256, 48000, stream=s), 1, 256)
h1 = bench.synth_irs(0, 64, 0, 240000)
probe("configs[1] twostage x64 head 128 L=240000", lambda s: F.TwoStageFFTConvolver.init(h1, 128, 240000, stream=s, async_tail=True), 64, 128)
h2 = bench.synth_irs(0, 256, 0, 96000)
probe("configs[2] crossfade x256 B=512 L=96000", lambda s: F.CrossfadeConvolver.init(h2, 512, 96000, stream=s), 256, 512)
