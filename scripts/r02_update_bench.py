#!/usr/bin/env python
"""Worst block time while an impulse-response update is in flight, 256 and 4096 channels x 2 s IRs x
block 512, end-to-end host path, one block per call.  Three ways to update every channel's response:
  sync        fcb_fftconv_update (the reference's semantics: returns when the new spectra are queued; pageable source)
  begin+wait  fcb_fftconv_update_begin(FCB_UPDATE_WAIT) from page-locked memory: returns at once, the next block waits
              ON THE DEVICE for K5 if it is still running (same output as `sync`)
  begin       fcb_fftconv_update_begin from page-locked memory: returns at once, the new response is swapped in by the
              first block that finds K5 finished — no block ever waits
One JSON line per (channels, mode): time of the update call itself, block times around it, blocks until the flip."""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
import fft_convolution_b200 as F

B, L = 512, 96000


def pinned(lib, shape):
    n = int(np.prod(shape)) * 4
    p = lib.fcb_host_alloc(n)
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=shape), p


def run(Cn, mode):
    lib = F.load_library()
    base = bench.synth_irs(0, min(Cn, 512), 0, L)
    reps = (Cn + base.shape[0] - 1) // base.shape[0]
    h0 = np.tile(base, (reps, 1))[:Cn]
    conv = F.FFTConvolver.init(h0, B, L)
    conv.update_reserve()
    hp_view, hp = pinned(lib, (Cn, L))
    hp_view[...] = h0[::-1]
    h_page = np.ascontiguousarray(h0[::-1])
    xin, pin = pinned(lib, (Cn, B))
    yout, pout = pinned(lib, (Cn, B))
    xin[...] = bench.synth_noise(0, Cn, 0, B)
    from fft_convolution_b200 import _lib

    def block():
        t0 = time.perf_counter()
        _lib.check(lib.fcb_fftconv_process(conv._h, pin, B, B, pout, B, B))
        return (time.perf_counter() - t0) * 1e3

    for _ in range(50):
        block()
    quiet = [block() for _ in range(200)]
    upd_ms, during, flips = [], [], []
    for rep in range(5):
        t0 = time.perf_counter()
        if mode == "sync":
            conv.update(h_page)
        else:
            conv.update_begin(hp, L, wait=(mode == "begin+wait"))
        upd_ms.append((time.perf_counter() - t0) * 1e3)
        n = 0
        for i in range(60):
            during.append(block())
            if conv.update_pending():
                n = i + 1
        flips.append(n)
        for _ in range(20):
            block()
    out = {"channels": Cn, "mode": mode, "block_period_ms": 1000.0 * B / 48000, "ir_bytes_MB": Cn * L * 4 / 1e6,
           "update_call_ms_mean": float(np.mean(upd_ms)), "update_call_ms_max": float(np.max(upd_ms)),
           "block_ms_quiet_p50": float(np.median(quiet)), "block_ms_quiet_max": float(np.max(quiet)),
           "block_ms_during_update_max": float(np.max(during)), "block_ms_during_update_p50": float(np.median(during)),
           "blocks_until_new_response_plays": flips}
    print(json.dumps(out), flush=True)
    conv.close()
    for p in (hp, pin, pout):
        lib.fcb_host_free(p)


if __name__ == "__main__":
    for Cn in (256, 4096):
        for mode in ("sync", "begin+wait", "begin"):
            run(Cn, mode)
