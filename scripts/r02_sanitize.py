#!/usr/bin/env python
"""Small runs of every round-2 kernel path for compute-sanitizer (memcheck / racecheck): split whole-block kernels (mono,
pair), bulk-I/O pair kernel on mapped staging, wide-plan transforms (B = 4096), background update + commit, nested partition."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
import fft_convolution_b200 as F

def run(conv, C, B, n, upd=None):
    out = np.zeros((C, B), np.float32) if C > 1 else np.zeros(B, np.float32)
    for i in range(n):
        x = bench.synth_noise(0, C, i * B, B)
        if upd and i == upd[0]:
            conv.update(upd[1])
        conv.process(x if C > 1 else x[0], out)
    return float(np.abs(out).max())

h = bench.synth_irs(0, 6, 0, 6000)
print("uniform split", run(F.FFTConvolver.init(h[0], 256, 6000), 1, 256, 30))
print("uniform batch", run(F.FFTConvolver.init(h, 128, 6000), 6, 128, 30))
print("twostage pair", run(F.TwoStageFFTConvolver.init(h, 64, 6000, async_tail=True), 6, 64, 80))
print("crossfade pair", run(F.CrossfadeConvolver.new(F.FFTConvolver.init(h, 128, 6000), 6000, 128, 300), 6, 128, 30, upd=(5, h[::-1].copy())))
hl = bench.synth_irs(0, 2, 0, 4096 * 3 + 7)
print("wide plan B=4096", run(F.FFTConvolver.init(hl, 4096, 4096 * 3 + 7), 2, 4096, 5))
print("nested", run(F.TwoStageFFTConvolver.init(h[0], 32, 6000, stages=3), 1, 32, 200))
g = F.FFTConvolver.init(h, 128, 6000)
g.update_reserve()
lib = F.load_library()
import ctypes as C
p = lib.fcb_host_alloc(h.nbytes)
v = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=h.shape)
v[...] = h[::-1]
out = np.zeros((6, 128), np.float32)
for i in range(20):
    if i == 4:
        g.update_begin(p, 6000, wait=True)
    if i == 9:
        g.update_begin(p, 6000)
    g.process(bench.synth_noise(0, 6, i * 128, 128), out)
g.sync()
lib.fcb_host_free(p)
print("background update ok")
