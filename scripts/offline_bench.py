#!/usr/bin/env python
"""Multi-block calls (time-batched pass) on one GPU, device-resident buffers, CUDA-event timing:
 (a) the headline shape — 4096 independent channels x 2 s IR x block 512 — with 1, 2, 4, 8 blocks per call;
 (b) the reference's example shape (examples/compare_partitioned.rs: mono, 64-sample blocks, 128 000-tap IR,
     1000 blocks) block by block vs one call.
One JSON line each.  Not bench.py lines: the headline is one block per call (the real-time case)."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import fft_convolution_b200 as F  # noqa: E402

SR = 48000


def main():
    C, B, L = 4096, 512, 96000
    st = torch.cuda.Stream()
    conv = F.FFTConvolver.init(bench.synth_irs(0, C, 0, L), B, L, stream=st.cuda_stream)
    conv.reserve(B * 16)  # process() never allocates: the multi-block workspace is sized here
    import os
    from fft_convolution_b200 import _lib
    for kv in filter(None, os.environ.get("FCB_TUNE", "").split(",")):
        k, v = kv.split("=")
        _lib.check(_lib.load().fcb_tune(k.encode(), int(v)))
    for nb in (1, 2, 4, 8, 16):
        x = torch.from_numpy(bench.synth_noise(0, C, 0, B * nb)).cuda()
        out = torch.empty((C, B * nb), dtype=torch.float32, device="cuda")
        n = B * nb
        for _ in range(5):
            conv.process_dev(x.data_ptr(), n, n, out.data_ptr(), n, n)
        torch.cuda.synchronize()
        steps = max(20, 160 // nb)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps):
            conv.process_dev(x.data_ptr(), n, n, out.data_ptr(), n, n)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(json.dumps({"config": f"FFTConvolver x{C} channels, 2 s IR, block {B}, {nb} block(s) per call (device buffers)",
                          "ms_per_call": ms, "ms_per_block": ms / nb, "channel_sec_per_sec": C * n / SR / (ms / 1e3)}), flush=True)
    # the same from pinned host buffers (H2D + pass + D2H per call; channel groups overlap copies and passes)
    nb = 16
    hx = torch.from_numpy(bench.synth_noise(0, C, 0, B * nb)).pin_memory()
    hy = torch.empty((C, B * nb), dtype=torch.float32).pin_memory()
    xin, yout = hx.numpy(), hy.numpy()
    conv.process(xin, yout)
    t0 = time.perf_counter()
    for _ in range(5):
        conv.process(xin, yout)
    ms = (time.perf_counter() - t0) / 5 * 1e3
    print(json.dumps({"config": f"FFTConvolver x{C} channels, 2 s IR, block {B}, {nb} blocks per call, pinned HOST buffers",
                      "ms_per_call": ms, "ms_per_block": ms / nb, "channel_sec_per_sec": C * B * nb / SR / (ms / 1e3),
                      "h2d_plus_d2h_MB": 2 * C * B * nb * 4 / 1e6}), flush=True)
    conv.close()
    # (b) the reference example's shape, mono
    B, L, nblocks = 64, 128000, 1000
    h = bench.synth_irs(0, 1, 0, L)[0]
    x = bench.synth_noise(0, 1, 0, B * nblocks)[0]
    g = F.FFTConvolver.init(h, B, L)
    g.reserve(x.size)  # process() never allocates: the multi-block workspace is sized here
    y1, y2 = np.zeros_like(x), np.zeros_like(x)
    blk = np.zeros(B, np.float32)
    g.process(x, y2)  # warm-up
    g.reset()
    t0 = time.perf_counter()
    for b in range(nblocks):
        g.process(x[b * B:(b + 1) * B], blk)
        y1[b * B:(b + 1) * B] = blk
    t_blocks = time.perf_counter() - t0
    g.reset()
    t0 = time.perf_counter()
    g.process(x, y2)
    t_call = time.perf_counter() - t0
    print(json.dumps({"config": f"mono FFTConvolver, block {B}, {L}-tap IR, {nblocks} blocks (reference example shape), host buffers",
                      "block_by_block_ms": t_blocks * 1e3, "one_call_ms": t_call * 1e3,
                      "max_abs_diff_over_rms": float(np.max(np.abs(y1 - y2))) / float(np.sqrt(np.mean(y2.astype(np.float64) ** 2))),
                      "audio_seconds": B * nblocks / SR}), flush=True)


if __name__ == "__main__":
    main()
