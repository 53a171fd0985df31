#!/usr/bin/env python
"""Sweep of the split target for mid-size batches (device buffers, CUDA events): channels x split_slots -> ms per block."""
import json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
import fft_convolution_b200 as F
from fft_convolution_b200 import _lib
lib = _lib.load()
B, L = 512, 96000
for C in (256, 512, 768, 1024):
    irs = bench.synth_irs(0, C, 0, L)
    x = torch.from_numpy(bench.synth_noise(0, C, 0, B)).cuda()
    y = torch.empty((C, B), device="cuda")
    for slots in (1, 592, 1184, 1776, 2368):
        _lib.check(lib.fcb_tune(b"split_slots", slots))
        st = torch.cuda.Stream()
        conv = F.FFTConvolver.init(irs, B, L, stream=st.cuda_stream)
        for _ in range(20):
            conv.process_dev(x.data_ptr(), B, B, y.data_ptr(), B, B)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(300):
            conv.process_dev(x.data_ptr(), B, B, y.data_ptr(), B, B)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 300
        print(json.dumps({"channels": C, "split_slots": slots, "ms_per_block": ms, "GBs": C * 16 * 188 * 513 / ms / 1e6}), flush=True)
        conv.close()
