"""Multi-block calls (offline_kernels.cuh): a process() call spanning several whole blocks runs as one
time-batched pass (sliding window of input spectra in registers).  The contract is bit-identity with the
block-by-block path — same operations in the same order — plus parity with the CPU oracle."""
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


def _run_calls(F, h, B, L, x, calls, multi):
    _tune(b"multi_block", 1 if multi else 0)
    try:
        g = F.FFTConvolver.init(h, B, L)
        y = np.zeros_like(x)
        pos = 0
        for n in calls:
            out = np.zeros(x.shape[:-1] + (n,), np.float32)
            g.process(np.ascontiguousarray(x[..., pos:pos + n]), out)
            y[..., pos:pos + n] = out
            pos += n
        return y[..., :pos]
    finally:
        _tune(b"multi_block", 1)


@pytest.mark.parametrize("C,B,L,calls", [
    (3, 64, 64 * 9 + 5, [64 * 2, 64 * 5, 64, 64 * 3 + 17, 47, 64 * 4, 64 * 7]),     # ragged tails, mid-block starts
    (2, 128, 128 * 3 + 1, [128 * 9, 128 * 2, 128 * 6]),                              # more blocks per call than segments
    (1, 16, 16 * 40, [16 * 64, 16 * 3, 16 * 33]),                                    # small blocks, long calls
    (5, 512, 512 * 12 + 100, [512 * 4, 512 * 3, 512 * 2, 512, 512 * 5]),             # the headline block size
    (2, 1024, 1024 * 5, [1024 * 3, 1024 * 6]),                                       # above the fused-kernel range
    (1100, 32, 32 * 6 + 1, [32 * 5, 32 * 2, 32 * 3]),                                # >= 1024 channels: grouped, copies overlap the passes
])
def test_multi_block_calls_are_bit_identical_to_block_by_block(F, C, B, L, calls):
    h = np.stack([oracle.gen_ir(c, 0, L) for c in range(C)])
    n = sum(calls)
    x = np.stack([oracle.gen_noise(40 + c, 0, n) for c in range(C)])
    a = _run_calls(F, h, B, L, x, calls, multi=True)
    b = _run_calls(F, h, B, L, x, calls, multi=False)
    assert np.array_equal(a, b)
    # and the same signal cut block by block: identical bits when every call was whole blocks, f32 rounding
    # otherwise (a partially filled block is transformed as it stands, src/fft_convolver.rs:234-241)
    blocks = [B] * (n // B) + ([n % B] if n % B else [])
    c = _run_calls(F, h, B, L, x, blocks, multi=False)
    if all(k % B == 0 for k in calls):
        assert np.array_equal(a, c)
    else:
        assert np.max(np.abs(a - c)) <= 1e-5 * rms(c)


def test_multi_block_mono_matches_oracle(F):
    B, L, n = 64, 3000, 64 * 150 + 13
    h = oracle.gen_ir(7, 0, L)
    x = oracle.gen_noise(8, 0, n)
    g, o = F.FFTConvolver.init(h, B, L), oracle.FFTConvolver.init(h, B, L)
    y, ref = np.zeros(n, np.float32), np.zeros(n, np.float32)
    g.process(x, y)   # one call: 150 whole blocks time-batched + a 13-sample tail
    o.process(x, ref)
    assert np.max(np.abs(y - ref)) <= 1e-5 * rms(ref)


def test_multi_block_shared_ir_and_state_carry_over(F):
    """channels sharing one IR; after a multi-block call the ring, overlap and `current` are what
    block-by-block processing leaves behind (the following single-block calls agree bit for bit)"""
    C, B, L = 6, 256, 256 * 7 + 3
    h = oracle.gen_ir(3, 0, L)
    x = np.stack([oracle.gen_noise(70 + c, 0, B * 12) for c in range(C)])
    outs = []
    for multi in (True, False):
        _tune(b"multi_block", 1 if multi else 0)
        g = F.FFTConvolver.init(h, B, L, channels=C)
        y = np.zeros_like(x)
        first = np.zeros((C, B * 8), np.float32)
        g.process(np.ascontiguousarray(x[:, :B * 8]), first)
        y[:, :B * 8] = first
        for b in range(8, 12):
            blk = np.zeros((C, B), np.float32)
            g.process(np.ascontiguousarray(x[:, b * B:(b + 1) * B]), blk)
            y[:, b * B:(b + 1) * B] = blk
        outs.append(y)
    _tune(b"multi_block", 1)
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("B,buf,fade", [(64, 256, 300), (32, 32 * 5, 40)])
def test_crossfade_with_a_buffer_of_several_blocks(F, B, buf, fade):
    """CrossfadeConvolver::new(conv, L, max_buffer_size = several blocks, fade): both inner convolvers take the
    multi-block path, the second with the crossfade-mix epilogue over the whole call; vs the oracle incl. updates"""
    L = B * 7 + 3
    h = oracle.gen_ir(1, 0, L)
    g = F.CrossfadeConvolver.new(F.FFTConvolver.init(h, B, L), L, buf, fade)
    o = oracle.CrossfadeConvolver.new(oracle.FFTConvolver.init(h, B, L), L, buf, fade)
    scale = 0.05
    for step in range(24):
        if step in (3, 4, 11, 17):
            hn = oracle.gen_ir(1, step, L if step != 11 else L // 2)
            g.update(hn)
            o.update(hn)
        x = oracle.gen_noise(5, step * buf, buf)
        yg, yo = np.zeros(buf, np.float32), np.zeros(buf, np.float32)
        g.process(x, yg)
        o.process(x, yo)
        scale = max(scale, rms(yo))
        assert np.max(np.abs(yg - yo)) <= 1e-5 * scale, step
        cnt, mix, appr, tgt = g.state()
        s = o.crossfader
        assert (cnt, appr, tgt, np.float32(mix)) == (s.counter, bool(s.approaching), s.target, np.float32(s.mix_value))
