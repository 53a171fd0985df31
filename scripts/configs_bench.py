#!/usr/bin/env python
"""Absolute numbers for the BASELINE configs that are parity cases rather than bench.py lines
(configs[0..2]): wall time per host-API call (H2D + kernels + D2H + sync) on one GPU, with the
CPU restatement beside it for config 0.  One JSON line per config."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import fft_convolution_b200 as F  # noqa: E402

SR = 48000


def timed(fn, n, warm=20):
    for i in range(warm):
        fn(i)
    t0 = time.perf_counter()
    for i in range(n):
        fn(warm + i)
    return (time.perf_counter() - t0) / n


def cfg0():
    B, L = 256, 48000
    h = bench.synth_irs(0, 1, 0, L)[0]
    x = bench.synth_noise(0, 1, 0, B * 1900)[0]
    conv = F.FFTConvolver.init(h, B, L)
    out = np.zeros(B, np.float32)
    dt = timed(lambda i: conv.process(x[(i % 1800) * B:(i % 1800 + 1) * B], out), 1500)
    import oracle  # CPU restatement beside it (baseline only)
    o = oracle.FFTConvolver.init(h, B, L)
    dto = timed(lambda i: o.process(x[(i % 1800) * B:(i % 1800 + 1) * B], out), 1500)
    return {"config": "configs[0] FFTConvolver mono, block 256, IR 48000", "us_per_block_gpu_host_api": dt * 1e6,
            "us_per_block_cpu_port_1core": dto * 1e6, "block_period_us": 1e6 * B / SR,
            "note": "one channel, state fits L2: one launch per block, the delay line cut over ~20 CTAs (fcb_tune split), I/O through "
                    "mapped pinned staging"}


def cfg1(async_tail, forced=0):
    C, H, L = 64, 128, 240000
    h = bench.synth_irs(0, C, 0, L)
    x = bench.synth_noise(0, C, 0, H * 64)
    conv = F.TwoStageFFTConvolver.init(h, H, L, async_tail=async_tail, forced_tail_block=forced)
    out = np.zeros((C, H), np.float32)
    blocks = [np.ascontiguousarray(x[:, i * H:(i + 1) * H]) for i in range(64)]
    n = 64 * 8  # spans 4 tail periods of T=8192 (64 head blocks each)
    per = []
    for i in range(64 + n):
        t0 = time.perf_counter()
        conv.process(blocks[i % 64], out)
        if i >= 64:
            per.append(time.perf_counter() - t0)
    per = np.array(per)
    import oracle  # CPU restatement beside it (baseline only): all host threads over the 64 channels
    xo = bench.synth_noise(0, C, 0, H * 256)
    t0 = time.perf_counter()
    oracle.batch_twostage(h, H, xo, H, forced_tail=forced)
    cpu_us = (time.perf_counter() - t0) / 256 * 1e6
    return {"config": f"configs[1] TwoStage x{C} ch, head 128, IR 240000, T={conv.tail_block_size}, async_tail={async_tail}",
            "us_per_head_block_cpu_port_all_threads": cpu_us, "cpu_threads": oracle.load().lib.orc_max_threads(),
            "us_per_head_block_mean": float(per.mean() * 1e6), "us_per_head_block_max": float(per.max() * 1e6),
            "us_per_head_block_p99": float(np.percentile(per, 99) * 1e6), "block_period_us": 1e6 * H / SR,
            "channel_sec_per_sec": C * H / SR / float(per.mean())}


def cfg2():
    C, B, L = 256, 512, 96000
    conv = F.CrossfadeConvolver.init(bench.synth_irs(0, C, 0, L), B, L)
    x = [bench.synth_noise(0, C, B * i, B) for i in range(8)]
    out = np.zeros((C, B), np.float32)
    upd = [bench.synth_irs(0, C, u, L) for u in (1, 2)]
    per, per_upd = [], []
    for i in range(260):
        if i and i % 50 == 0:
            t0 = time.perf_counter()
            conv.update(upd[(i // 50) % 2])
            conv.sync()
            per_upd.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        conv.process(x[i % 8], out)
        if i >= 10:
            per.append(time.perf_counter() - t0)
    per = np.array(per)
    import oracle  # CPU restatement beside it (baseline only)
    h0 = bench.synth_irs(0, C, 0, L)
    xo = bench.synth_noise(0, C, 0, B * 100)
    t0 = time.perf_counter()
    oracle.batch_crossfade(h0, B, xo, irs_upd=np.stack(upd), update_every=50)
    cpu_us = (time.perf_counter() - t0) / 100 * 1e6
    return {"config": f"configs[2] CrossfadeConvolver::init x{C} ch, block 512, IR 96000, update every 50 blocks",
            "us_per_block_cpu_port_all_threads": cpu_us, "cpu_threads": oracle.load().lib.orc_max_threads(),
            "us_per_block_mean": float(per.mean() * 1e6), "us_per_block_max": float(per.max() * 1e6),
            "ms_per_update_call": float(np.mean(per_upd) * 1e3), "block_period_us": 1e6 * B / SR,
            "channel_sec_per_sec": C * B / SR / float(per.mean()),
            "note": "both convolvers run every block like the reference (2x the FFTConvolver bytes)"}


if __name__ == "__main__":
    for fn in (cfg0, lambda: cfg1(False), lambda: cfg1(True), lambda: cfg1(True, 4096), cfg2):
        print(json.dumps(fn()), flush=True)
