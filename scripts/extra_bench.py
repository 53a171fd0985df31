#!/usr/bin/env python
"""Extra absolute numbers on one GPU (device-resident inputs, CUDA-event timing):
 (a) 4096 channels sharing ONE 2 s IR (shared_ir: spectra stored once, IR tiles hit in L2),
 (b) 4096 independent channels through TwoStageFFTConvolver (head 512 -> T = 8192) — the
     reference's non-uniform partition moves ~4.8x fewer bytes per block than the uniform one.
One JSON line each.  Not bench.py lines (the headline stays the uniform FFTConvolver config)."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import fft_convolution_b200 as F  # noqa: E402

SR, C, B, L = 48000, 4096, 512, 96000


def timed(step, steps=200, warm=20, stream=None):
    for i in range(warm):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        step(warm + i)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else "all"
    steps_ts = int(sys.argv[2]) if len(sys.argv) > 2 else 320
    import os
    from fft_convolution_b200 import _lib
    for kv in filter(None, os.environ.get("FCB_TUNE", "").split(",")):
        k, v = kv.split("=")
        _lib.check(_lib.load().fcb_tune(k.encode(), int(v)))
    st = torch.cuda.Stream()
    x = [torch.from_numpy(bench.synth_noise(0, C, B * i, B)).cuda() for i in range(8)]
    out = torch.empty((C, B), dtype=torch.float32, device="cuda")
    # (a) shared IR
    if only in ("all", "shared"):
      conv = F.FFTConvolver.init(bench.synth_irs(0, 1, 0, L)[0], B, L, channels=C, stream=st.cuda_stream)
      ms = timed(lambda i: conv.process_dev(x[i % 8].data_ptr(), B, B, out.data_ptr(), B, B), stream=st)
      S, K = conv.seg_count, B + 1
      print(json.dumps({"config": f"FFTConvolver x{C} channels sharing one IR (2 s, block 512)", "ms_per_block": ms,
                        "channel_sec_per_sec": C * B / SR / (ms / 1e3),
                        "ring_GBs": C * 8 * (S - 1) * K / (ms / 1e3) / 1e9}), flush=True)
      conv.close()
    if only in ("all", "crossfade"):
        # (c) configs[2] at scale: CrossfadeConvolver x 4096 channels, fading all the time (update every 50 blocks, 2 s fade)
        irs = bench.synth_irs(0, C, 0, L)
        xf = F.CrossfadeConvolver.init(irs, B, L, stream=st.cuda_stream)
        upd = bench.synth_irs(0, C, 1, L)
        xf.update(upd)  # starts the 96 000-sample fade: A and B both audible from the second block on
        del irs, upd
        ms = timed(lambda i: xf.process_dev(x[i % 8].data_ptr(), B, B, out.data_ptr(), B, B), steps=100, warm=10, stream=st)
        S, K = 188, B + 1
        print(json.dumps({"config": f"CrossfadeConvolver::init x{C} channels, 2 s IRs, block 512, mid-fade (A + B in one paired launch)",
                          "ms_per_block": ms, "channel_sec_per_sec": C * B / SR / (ms / 1e3),
                          "GBs_of_24SK_bytes": C * 24 * (S - 1) * K / (ms / 1e3) / 1e9}), flush=True)
        xf.close()
    if only not in ("all", "twostage"):
        return
    # (b) two-stage at scale
    irs = bench.synth_irs(0, C, 0, L)
    ts = F.TwoStageFFTConvolver.init(irs, B, L, stream=st.cuda_stream, async_tail=True)
    del irs
    ms = timed(lambda i: ts.process_dev(x[i % 8].data_ptr(), B, B, out.data_ptr(), B, B), steps=steps_ts, warm=64, stream=st)
    ts.sync()
    print(json.dumps({"config": f"TwoStageFFTConvolver x{C} independent channels, head 512, T={ts.tail_block_size}, IR 2 s, async tail",
                      "ms_per_block_mean": ms, "channel_sec_per_sec": C * B / SR / (ms / 1e3)}), flush=True)


if __name__ == "__main__":
    main()
