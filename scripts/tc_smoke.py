"""One small tensor-core MIMO run against the CUDA-core path (debug helper for K4)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import fft_convolution_b200 as F

n_out, n_in, B, L, NS, nblocks = 16, 2, 64, 64 * 20 + 5, 3, int(sys.argv[1]) if len(sys.argv) > 1 else 25
rng = np.random.default_rng(1)
h = (rng.standard_normal((n_out, n_in, L)) * np.exp(-np.arange(L) / (L / 6.9))).astype(np.float32)
x = (rng.random((NS * n_in, B * nblocks), dtype=np.float32) * 2 - 1).astype(np.float32)
ys = {}
for tc in (False, True):
    g = F.MimoConvolver.init(h, B, L, n_streams=NS, tensor_cores=tc)
    print("tensor cores:", g.uses_tensor_cores, flush=True)
    y = np.zeros((NS * n_out, B * nblocks), np.float32)
    o = np.zeros((NS * n_out, B), np.float32)
    for b in range(nblocks):
        g.process(np.ascontiguousarray(x[:, b * B:(b + 1) * B]), o)
        y[:, b * B:(b + 1) * B] = o
    ys[tc] = y
r = np.sqrt(np.mean(ys[False] ** 2, axis=1))
e = np.max(np.abs(ys[True] - ys[False]), axis=1) / r
print("max err / rms per row:", e.max(), "rows worst:", np.argsort(e)[-4:], "rms", r[:3])
for b in range(nblocks):
    eb = np.max(np.abs(ys[True][:, b * B:(b + 1) * B] - ys[False][:, b * B:(b + 1) * B]))
    print(b, eb)
