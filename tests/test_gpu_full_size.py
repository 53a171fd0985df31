"""Parity at BASELINE.json's full sizes (reduced only in channel count / run length where the
check is size-independent): every run is long enough for the segment ring to wrap, outputs are
compared with an f64 FFT convolution of the same synthetic inputs (<= 1e-5 * output RMS), and the
linearity / shift properties of the operator are checked on the device path itself."""
import numpy as np
import pytest

import bench
from refsignals import rms

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def F():
    import fft_convolution_b200 as f
    return f


def truth(x, h):
    n = x.shape[-1]
    nfft = 1 << int(np.ceil(np.log2(n + h.shape[-1])))
    return np.fft.irfft(np.fft.rfft(x.astype(np.float64), nfft) * np.fft.rfft(h.astype(np.float64), nfft), nfft)[..., :n]


def run_blocks(conv, x, B):
    y = np.zeros_like(x)
    blk = np.zeros(x.shape[:-1] + (B,), np.float32)
    for b in range(x.shape[-1] // B):
        conv.process(np.ascontiguousarray(x[..., b * B:(b + 1) * B]), blk)
        y[..., b * B:(b + 1) * B] = blk
    return y


def test_config3_4_shape_full_ir_ring_wraps(F):
    """configs[3]: 2 s IR (96 000 taps, S = 188), block 512 — 24 channels, 260 blocks (> S)"""
    C, B, L, NB = 24, 512, 96000, 260
    h = bench.synth_irs(0, C, 0, L)
    x = bench.synth_noise(0, C, 0, B * NB)
    conv = F.FFTConvolver.init(h, B, L)
    assert conv.seg_count == 188
    y = run_blocks(conv, x, B)
    t = truth(x, h)
    for c in range(C):
        assert np.max(np.abs(y[c] - t[c])) <= TOL * rms(t[c]), c
    assert conv.current == (188 - NB % 188) % 188


def test_config3_4_linearity_and_shift_on_device(F):
    """size-independent properties of the operator at the full IR length"""
    C, B, L, NB = 3, 512, 96000, 40
    h = bench.synth_irs(7, 1, 0, L)[0]
    x1, x2 = bench.synth_noise(1, 1, 0, B * NB)[0], bench.synth_noise(2, 1, 0, B * NB)[0]
    a, b = np.float32(0.75), np.float32(-1.5)
    shifted = np.concatenate([np.zeros(B, np.float32), x1[:-B]])  # delayed by exactly one block
    xs = np.stack([x1, x2, (a * x1 + b * x2).astype(np.float32), shifted])
    y = run_blocks(F.FFTConvolver.init(h, B, L, channels=4), xs, B)  # shared IR, 4 channels
    r = rms(y[2])
    assert np.max(np.abs(y[2] - (a * y[0] + b * y[1]))) <= 4 * TOL * r            # linearity
    assert np.max(np.abs(y[3][B:] - y[0][:-B])) <= 4 * TOL * rms(y[0]) and np.all(y[3][:B] == 0)  # time invariance


def test_config0_shape_full_run_prefix(F):
    """configs[0]: mono, block 256, 48 000-tap IR, white noise — 400 blocks (> S = 188)"""
    B, L, NB = 256, 48000, 400
    h = bench.synth_irs(0, 1, 0, L)[0]
    x = bench.synth_noise(0, 1, 0, B * NB)[0]
    y = run_blocks(F.FFTConvolver.init(h, B, L), x, B)
    t = truth(x, h)
    assert np.max(np.abs(y - t)) <= TOL * rms(t)


def test_config1_shape_twostage_full_ir(F):
    """configs[1]: head 128, 5 s IR (240 000 taps) -> T = 8192; 4 channels, 3 tail periods"""
    C, H, L = 4, 128, 240000
    h = bench.synth_irs(0, C, 0, L)
    x = bench.synth_noise(0, C, 0, H * 64 * 3 + H * 5)
    conv = F.TwoStageFFTConvolver.init(h, H, L, async_tail=True)
    assert conv.tail_block_size == 8192
    n = (x.shape[1] // H) * H
    y = run_blocks(conv, x[:, :n], H)
    t = truth(x[:, :n], h)
    for c in range(C):
        assert np.max(np.abs(y[c] - t[c])) <= TOL * rms(t[c]), c


def test_config4_shape_mimo_full_matrix(F):
    """configs[4]: 16 x 16 matrix, 10 s IRs (480 000 taps, S = 938), block 512; 960 blocks so the
    ring wraps; two outputs checked against the f64 sum of 16 convolutions"""
    N, B, L, NB = 16, 512, 480000, 960
    h = bench.synth_irs(0, N * N, 0, L).reshape(N, N, L)
    x = bench.synth_noise(0, N, 0, B * NB)
    m = F.MimoConvolver.init(h, B, L)
    assert m.seg_count == 938
    y = np.zeros((N, B * NB), np.float32)
    blk = np.zeros((N, B), np.float32)
    for b in range(NB):
        m.process(np.ascontiguousarray(x[:, b * B:(b + 1) * B]), blk)
        y[:, b * B:(b + 1) * B] = blk
    for o in (0, 11):
        t = np.zeros(B * NB)
        for i in range(N):
            t += truth(x[i], h[o, i])
        assert np.max(np.abs(y[o] - t)) <= TOL * rms(t), o


def test_config1_full_channel_count_vs_oracle(F):
    """configs[1] at its full 64 channels: TwoStage head 128 / T = 8192 (derived), 5 s IRs, 200 head blocks (three tail
    periods and a bit) — every channel against its own oracle convolver (OpenMP over channels on the CPU side)"""
    import oracle
    C, H, L, NB = 64, 128, 240000, 200
    h = bench.synth_irs(0, C, 0, L)
    x = bench.synth_noise(0, C, 0, H * NB)
    conv = F.TwoStageFFTConvolver.init(h, H, L, async_tail=True)
    assert conv.tail_block_size == 8192
    y = run_blocks(conv, x, H)
    ref = oracle.batch_twostage(h, H, x, H)
    worst = max(float(np.max(np.abs(y[c] - ref[c]))) / rms(ref[c]) for c in range(C))
    assert worst <= TOL, f"configs[1] x 64 channels: {worst:.3e} x RMS"


def test_config2_full_channel_count_vs_oracle(F):
    """configs[2] at its full 256 channels: CrossfadeConvolver::init(h, 512, 96 000) — a 96 000-sample fade after a
    512-sample hold — with update() every 50 blocks, 120 blocks, every channel against its own oracle convolver"""
    import oracle
    C, B, L, NB = 256, 512, 96000, 120
    h = bench.synth_irs(0, C, 0, L)
    upd = np.stack([bench.synth_irs(0, C, u, L) for u in (1, 2)])
    x = bench.synth_noise(0, C, 0, B * NB)
    conv = F.CrossfadeConvolver.init(h, B, L)
    y = np.zeros_like(x)
    blk = np.zeros((C, B), np.float32)
    for b in range(NB):
        if b and b % 50 == 0:
            conv.update(upd[(b // 50 - 1) % 2])
        conv.process(np.ascontiguousarray(x[:, b * B:(b + 1) * B]), blk)
        y[:, b * B:(b + 1) * B] = blk
    ref = oracle.batch_crossfade(h, B, x, irs_upd=upd, update_every=50)
    worst = max(float(np.max(np.abs(y[c] - ref[c]))) / rms(ref[c]) for c in range(C))
    assert worst <= TOL, f"configs[2] x 256 channels: {worst:.3e} x RMS"


def test_config3_full_channel_count_through_the_headline_path(F):
    """configs[3] at its FULL size: 4096 channels x 2 s IR x block 512, 190 blocks (the 188-slot ring wraps), through the
    very path bench.py times end to end — fcb_fftconv_process on page-locked host buffers (the whole-block kernel pulls and
    pushes the blocks itself) — every channel against its own oracle convolver (OpenMP over channels)."""
    import ctypes as C
    import oracle
    from fft_convolution_b200 import _lib
    Cn, B, L, NB = 4096, 512, 96000, 190
    lib = F.load_library()
    h = bench.synth_irs(0, Cn, 0, L)                             # 4096 different responses
    x = np.tile(bench.synth_noise(0, 256, 0, B * NB), (Cn // 256, 1))  # 256 different inputs, each met by 16 responses
    conv = F.FFTConvolver.init(h, B, L)
    p_in, p_out = lib.fcb_host_alloc(Cn * B * 4), lib.fcb_host_alloc(Cn * B * 4)
    h_in = np.ctypeslib.as_array(C.cast(p_in, C.POINTER(C.c_float)), shape=(Cn, B))
    h_out = np.ctypeslib.as_array(C.cast(p_out, C.POINTER(C.c_float)), shape=(Cn, B))
    y = np.zeros_like(x)
    launches0 = lib.fcb_launch_count()
    for b in range(NB):
        h_in[...] = x[:, b * B:(b + 1) * B]
        _lib.check(lib.fcb_fftconv_process(conv._h, p_in, B, B, p_out, B, B))
        y[:, b * B:(b + 1) * B] = h_out
    assert lib.fcb_launch_count() - launches0 == NB  # one fused launch per block: the zero-copy path was taken
    assert conv.current == (188 - NB % 188) % 188
    ref = oracle.batch_fftconv(h, B, x, B)
    err = np.max(np.abs(y - ref), axis=1) / np.sqrt(np.mean(ref.astype(np.float64) ** 2, axis=1))
    assert float(err.max()) <= TOL, f"configs[3] x 4096 channels: worst channel {int(err.argmax())} at {float(err.max()):.3e} x RMS"
    lib.fcb_host_free(p_in)
    lib.fcb_host_free(p_out)
