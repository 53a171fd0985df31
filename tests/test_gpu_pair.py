"""Paired launch (k_block_fused_pair): two convolvers fed the same input — TwoStage's head + tail_convolver0,
Crossfade's A + B — share one forward FFT and one ring stream.  Contract: bit-identical to separate launches."""
import numpy as np
import pytest

import oracle
from refsignals import rms

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("exact_paths")]


@pytest.fixture(scope="module")
def F():
    import fft_convolution_b200 as f
    return f


def _tune(key, value):
    from fft_convolution_b200 import _lib
    _lib.check(_lib.load().fcb_tune(key, value))


@pytest.mark.parametrize("C,H,L,calls", [
    (3, 64, 64 * 80, [64] * 70),                                  # T = 1024: head / tail0 16 segments each, whole blocks
    (2, 128, 128 * 70 + 9, [128, 128, 50, 78, 128, 128, 17, 111] * 6),  # ragged calls in between: pairing only on whole blocks
    (5, 512, 512 * 40, [512] * 40),                               # head 512 -> T = 8192 (the at-scale shape, fewer channels)
])
def test_two_stage_pair_is_bit_identical(F, C, H, L, calls):
    h = np.stack([oracle.gen_ir(c, 0, L) for c in range(C)])
    n = sum(calls)
    x = np.stack([oracle.gen_noise(60 + c, 0, n) for c in range(C)])
    outs = {}
    for pair in (1, 0):
        _tune(b"fused_pair", pair)
        try:
            g = F.TwoStageFFTConvolver.init(h, H, L)
            y = np.zeros_like(x)
            pos = 0
            for k in calls:
                o = np.zeros((C, k), np.float32)
                g.process(np.ascontiguousarray(x[:, pos:pos + k]), o)
                y[:, pos:pos + k] = o
                pos += k
            outs[pair] = y
        finally:
            _tune(b"fused_pair", 1)
    assert np.array_equal(outs[1], outs[0])
    ref = np.zeros(n, np.float32)
    o = oracle.TwoStageFFTConvolver.init(h[0], H, L)
    pos = 0
    for k in calls:
        t = np.zeros(k, np.float32)
        o.process(x[0, pos:pos + k], t)
        ref[pos:pos + k] = t
        pos += k
    assert np.max(np.abs(outs[1][0] - ref)) <= 1e-5 * rms(ref)


def test_crossfade_pair_is_bit_identical(F):
    C, B, L, fade = 4, 256, 256 * 9 + 3, 700
    h = np.stack([oracle.gen_ir(c, 0, L) for c in range(C)])
    outs, states = {}, {}
    for pair in (1, 0):
        _tune(b"fused_pair", pair)
        try:
            g = F.CrossfadeConvolver.new(F.FFTConvolver.init(h, B, L), L, B, fade)
            y = []
            for step in range(40):
                if step in (5, 6, 19, 30):
                    g.update(np.stack([oracle.gen_ir(c, step, L if step != 19 else L - 300) for c in range(C)]))
                x = np.stack([oracle.gen_noise(90 + c, step * B, B) for c in range(C)])
                o = np.zeros((C, B), np.float32)
                g.process(x, o)
                y.append(o)
            outs[pair] = np.concatenate(y, axis=1)
            states[pair] = g.state()
        finally:
            _tune(b"fused_pair", 1)
    assert np.array_equal(outs[1], outs[0])
    assert states[1] == states[0]
