"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the CPU
oracle where the independent restatements agree): the oracle must reproduce them bit for bit
(CPU), the CUDA path to <= 1e-5 * output RMS (GPU)."""
from pathlib import Path

import numpy as np
import pytest

import oracle
from refsignals import rms

GOLDEN = sorted((Path(__file__).resolve().parent / "golden").glob("*.npz"))


def _replay(d, uniform, twostage, crossfade):
    kind = str(d["kind"])
    B, L, sizes = int(d["block"]), int(d["ir_len"]), [int(v) for v in d["sizes"]]
    x = d["x"]
    y = np.zeros_like(x)
    if kind == "crossfade":
        xf = crossfade(d["h"], L, B, int(d["fade"]))
        blk = np.zeros(B, np.float32)
        for i in range(x.size // B):
            if i == int(d["update_block"]):
                xf.update(d["h1"])
            xf.process(x[i * B:(i + 1) * B], blk)
            y[i * B:(i + 1) * B] = blk
        return y
    conv = (uniform if kind == "uniform" else twostage)(d["h"], B, L)
    p = k = 0
    while p < x.size:
        n = min(sizes[k % len(sizes)], x.size - p)
        blk = np.zeros(n, np.float32)
        conv.process(x[p:p + n], blk)
        y[p:p + n] = blk
        p += n
        k += 1
    return y


def test_golden_files_exist():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
def test_oracle_reproduces_golden(path):
    d = np.load(path)
    y = _replay(d, oracle.FFTConvolver.init, oracle.TwoStageFFTConvolver.init,
                lambda h, L, B, fade: oracle.CrossfadeConvolver.new(oracle.FFTConvolver.init(h, B, L), L, B, fade))
    assert np.array_equal(y, d["y"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
def test_engine_matches_golden(path):
    import fft_convolution_b200 as F
    d = np.load(path)
    y = _replay(d, F.FFTConvolver.init, F.TwoStageFFTConvolver.init,
                lambda h, L, B, fade: F.CrossfadeConvolver.new(F.FFTConvolver.init(h, B, L), L, B, fade))
    assert y.shape == d["y"].shape
    assert np.max(np.abs(y - d["y"])) <= 1e-5 * rms(d["y"])
