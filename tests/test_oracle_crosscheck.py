"""Oracle sanity beyond the reference's tests: C oracle vs the independent numpy
restatement vs f64 truth, chunk-size invariance, the tail-size table of SURVEY.md App. B,
and the synthetic-data generators.  CPU only."""
import numpy as np
import pytest

import oracle
from oracle import oracle_np
from refsignals import rms


def _run(conv, x, n):
    out = np.zeros_like(x)
    for p in range(0, x.size, n):
        m = min(n, x.size - p)
        blk = np.zeros(m, np.float32)
        conv.process(x[p:p + m], blk)
        out[p:p + m] = blk
    return out


@pytest.mark.parametrize("B,L,nx", [(256, 48000, 256 * 40), (64, 1000, 64 * 50), (512, 5000, 512 * 12)])
def test_c_vs_numpy_vs_truth(B, L, nx):
    h = oracle.gen_ir(0, 0, L)
    x = oracle.gen_noise(0, 0, nx)
    yc = _run(oracle.FFTConvolver.init(h, B, L), x, B)
    yn = _run(oracle_np.FFTConvolverNP.init(h, B, L), x, B)
    yt = oracle_np.truth_f64(x, h)
    r = rms(yt)
    assert np.max(np.abs(yc - yt)) <= 1e-5 * r
    assert np.max(np.abs(yn - yt)) <= 1e-5 * r
    assert np.max(np.abs(yc - yn)) <= 1e-5 * r


def test_direct_f64_agrees_with_fft_truth():
    h = oracle.gen_ir(3, 0, 300)
    x = oracle.gen_noise(3, 0, 2000)
    assert np.max(np.abs(oracle.direct_conv_f64(x, h) - oracle_np.truth_f64(x, h))) < 1e-12


def test_rfft_matches_numpy_f64():
    lib = oracle.load().lib
    rng = np.random.default_rng(1)
    for n in (2, 4, 8, 64, 512, 1024, 16384):
        p = lib.orc_plan_new(n)
        x = rng.standard_normal(n).astype(np.float32)
        X = np.zeros(n // 2 + 1, np.complex64)
        lib.orc_rfft_forward(p, x, X)
        ref = np.fft.rfft(x.astype(np.float64))
        assert np.max(np.abs(X - ref)) <= 2e-6 * np.sqrt(n) * np.max(np.abs(x))
        assert X[0].imag == 0 and X[-1].imag == 0
        y = np.zeros(n, np.float32)
        lib.orc_rfft_inverse(p, X, y)
        assert np.max(np.abs(y - x)) <= 2e-6 * np.max(np.abs(x)) * np.log2(n + 1)
        lib.orc_plan_free(p)


def test_chunk_size_invariance():
    """zero added latency / arbitrary call sizes (src/fft_convolver.rs:222-231): random
    chunk sizes 1..99 give the block-sized result up to f32 rounding."""
    B, L = 64, 1000
    h = oracle.gen_ir(1, 0, L)
    x = oracle.gen_noise(1, 0, 64 * 60)
    ref = _run(oracle.FFTConvolver.init(h, B, L), x, B)
    rng = np.random.default_rng(7)
    conv = oracle.FFTConvolver.init(h, B, L)
    out = np.zeros_like(x)
    p = 0
    while p < x.size:
        n = min(int(rng.integers(1, 100)), x.size - p)
        blk = np.zeros(n, np.float32)
        conv.process(x[p:p + n], blk)
        out[p:p + n] = blk
        p += n
    assert np.max(np.abs(out - ref)) <= 1e-5 * rms(ref)


def test_twostage_vs_truth_partial_calls():
    L, H = 12000, 64
    h = oracle.gen_ir(2, 0, L)
    x = oracle.gen_noise(2, 0, 64 * 300)
    yt = oracle_np.truth_f64(x, h)
    for n in (64, 48, 17):
        y = _run(oracle.TwoStageFFTConvolver.init(h, H, L), x, n)
        assert np.max(np.abs(y - yt)) <= 1e-5 * rms(yt)
    yn = _run(oracle_np.TwoStageNP.init(h, H, L), x, 64)
    assert np.max(np.abs(yn - yt)) <= 1e-5 * rms(yt)


def test_twostage_non_power_of_two_head_panics():
    """head 48, T = 512: tail_input overflows on the 11th block (src/fft_convolver.rs:459)"""
    h = oracle.gen_ir(0, 0, 5000)
    o = oracle.TwoStageFFTConvolver.init(h, 48, 5000)
    blk, out = np.zeros(48, np.float32), np.zeros(48, np.float32)
    for _ in range(10):
        o.process(blk, out)
    with pytest.raises(oracle.OraclePanic):
        o.process(blk, out)


@pytest.mark.parametrize("head,L,T", [(128, 240000, 8192), (128, 220500, 8192), (64, 128000, 4096),
                                      (64, 12000, 1024), (1024, 1024, 1024), (512, 96000, 8192),
                                      (256, 48000, 4096), (512, 480000, 16384), (128, 140002, 4096),
                                      (128, 140003, 8192)])
def test_tail_block_size_table(head, L, T):
    """src/fft_convolver.rs:514-526 evaluated in f32."""
    assert oracle.compute_tail_block_size(head, L) == T
    assert oracle_np.compute_tail_block_size(head, L) == T


def test_update_changes_segment_count_literal():
    """update() with a shorter IR re-interprets the ring modulo the new count
    (src/fft_convolver.rs:190, :248, :287-291); C and numpy restatements must agree."""
    B, L = 32, 320
    h0, h1 = oracle.gen_ir(5, 0, L), oracle.gen_ir(5, 1, 100)
    x = oracle.gen_noise(5, 0, 32 * 40)
    a, b = oracle.FFTConvolver.init(h0, B, L), oracle_np.FFTConvolverNP.init(h0, B, L)
    oa, ob = np.zeros(B, np.float32), np.zeros(B, np.float32)
    ya, yb = [], []
    for i in range(40):
        if i == 13:
            a.update(h1)
            b.update(h1)
            assert a.active_seg_count == b.active_seg_count == 4
        a.process(x[i * B:(i + 1) * B], oa)
        b.process(x[i * B:(i + 1) * B], ob)
        ya.append(oa.copy()); yb.append(ob.copy())
    ya, yb = np.concatenate(ya), np.concatenate(yb)
    assert np.max(np.abs(ya - yb)) <= 1e-5 * rms(ya)


def test_crossfader_accumulated_mix_value():
    """H3: 95 999 sequential f32 adds of -1/96000 end at -1.0009856, not -0.9999896."""
    x = oracle.Crossfader(96000, 0)
    x.fade_into(x.B)
    for _ in range(95999):
        x.mix(np.float32(0), np.float32(0))
    assert abs(x.s.mix_value - (-1.0009856)) < 2e-6


def test_generators():
    lib = oracle.load().lib
    assert lib.orc_mix64(0) == 0xE220A8397B1DCDAF  # splitmix64 first output for state 0
    x = oracle.gen_noise(3, 10, 1000)
    assert x.min() >= -1.0 and x.max() < 1.0 and abs(float(x.mean())) < 0.1
    assert np.array_equal(oracle.gen_noise(3, 0, 1010)[10:], x)
    h = oracle.gen_ir(2, 1, 4800)
    assert abs(float(np.sum(h.astype(np.float64) ** 2)) - 1.0) < 1e-6
    assert not np.array_equal(h, oracle.gen_ir(2, 2, 4800))


def test_batch_runner_matches_single():
    lib = oracle.load().lib
    C, B, L, calls = 3, 64, 500, 20
    irs = np.stack([oracle.gen_ir(c, 0, L) for c in range(C)])
    x = np.stack([oracle.gen_noise(c, 0, B * calls) for c in range(C)])
    out = np.zeros_like(x)
    t = lib.orc_batch_fftconv_run(C, B, L, irs, x, out, B, calls, 2)
    assert t > 0
    for c in range(C):
        assert np.array_equal(out[c], _run(oracle.FFTConvolver.init(irs[c], B, L), x[c], B))
