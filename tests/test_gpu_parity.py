"""Parity of the CUDA path against the CPU oracle and f64 truth on the synthetic workloads of
SURVEY.md §8(d): stage-level (K5/K1 spectra, K2 bit-exact, K3), whole-convolver at every block
size, ragged call sizes, update() with a changed segment count, batched channels, shared IR."""
import numpy as np
import pytest

import oracle
from oracle import oracle_np
from refsignals import rms, WholeRun

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("exact_paths")]
TOL = 1e-5


@pytest.fixture(scope="module")
def F():
    import fft_convolution_b200 as f
    return f


def _run(conv, x, sizes):
    out = np.zeros_like(x)
    p = 0
    k = 0
    while p < x.shape[-1]:
        n = min(sizes[k % len(sizes)], x.shape[-1] - p)
        blk = np.zeros(x.shape[:-1] + (n,), np.float32)
        conv.process(np.ascontiguousarray(x[..., p:p + n]), blk)
        out[..., p:p + n] = blk
        p += n
        k += 1
    return out


@pytest.mark.parametrize("B", [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384])
def test_every_block_size_vs_oracle_and_truth(F, B):
    L = max(3 * B + B // 2 + 1, 5)
    nblocks = 6 if B >= 2048 else 12
    h = oracle.gen_ir(B % 97, 0, L)
    x = oracle.gen_noise(B % 97, 0, B * nblocks)
    y = _run(F.FFTConvolver.init(h, B, L), x, [B])
    yo = _run(oracle.FFTConvolver.init(h, B, L), x, [B])
    yt = oracle_np.truth_f64(x, h)
    r = rms(yt)
    assert np.max(np.abs(y - yo)) <= TOL * r
    assert np.max(np.abs(y - yt)) <= TOL * r


def test_config1_shape_vs_oracle(F):
    """BASELINE configs[0]: mono, block 256, 48 000-tap IR (shortened run: 60 blocks)."""
    B, L = 256, 48000
    h = oracle.gen_ir(0, 0, L)
    x = oracle.gen_noise(0, 0, B * 60)
    conv = F.FFTConvolver.init(h, B, L)
    assert (conv.block_size, conv.seg_count) == (256, 188)
    y = _run(conv, x, [B])
    yo = _run(oracle.FFTConvolver.init(h, B, L), x, [B])
    yt = oracle_np.truth_f64(x, h)
    assert np.max(np.abs(y - yo)) <= TOL * rms(yt)
    assert np.max(np.abs(y - yt)) <= TOL * rms(yt)


def test_stage_spectra_and_bit_exact_mac(F):
    """K5/K1 spectra within f32 FFT noise of the oracle's; K2 fed the oracle's spectra is
    bit-identical to the reference loop (src/fft_convolver.rs:244-255); K3 within tolerance."""
    import ctypes as C
    from fft_convolution_b200 import _lib
    lib = _lib.load()
    B, L = 128, 128 * 9 + 17
    h = oracle.gen_ir(4, 0, L)
    x = oracle.gen_noise(4, 0, B * 14)
    g, o = F.FFTConvolver.init(h, B, L), oracle.FFTConvolver.init(h, B, L)
    S = o.seg_count
    assert g.seg_count == S == 10
    for i in range(S):  # K5
        ref = o.segment_ir(i)
        got = g.segment_ir(i)
        assert np.max(np.abs(got - ref)) <= 4e-6 * np.max(np.abs(ref)) + 1e-7
        assert got[0].imag == 0 and got[-1].imag == 0
    og, oo = np.zeros(B, np.float32), np.zeros(B, np.float32)
    for b in range(13):
        g.process(x[b * B:(b + 1) * B], og)
        o.process(x[b * B:(b + 1) * B], oo)
    assert g.current == o.current and g.fill == o.fill == 0
    for i in range(S):  # K1 ring contents
        ref = o.segment(i)
        assert np.max(np.abs(g.segment(i) - ref)) <= 4e-6 * np.max(np.abs(ref)) + 1e-7
    # K2 bit-exactness: load the oracle's exact spectra into the device, run K2 alone
    eng = g.engine
    for i in range(S):
        for fn, row in (("fcb_engine_write_ir_segment", o.segment_ir(i)), ("fcb_engine_write_ring_segment", o.segment(i))):
            buf = np.ascontiguousarray(row).view(np.float32)
            _lib.check(getattr(lib, fn)(eng, 0, i, buf.ctypes.data_as(C.c_void_p)))
    # one more block through the oracle so that its pre_multiplied is computed from exactly this ring
    cur = o.current
    o.process(x[13 * B:14 * B], oo)
    for impl in (1, 2):
        _lib.check(lib.fcb_tune(b"mac_impl", impl))
        _lib.check(lib.fcb_engine_mac(eng, cur, S))
        got = g.premul()
        ref = o.premul()
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), f"K2 impl {impl} not bit-exact"
    _lib.check(lib.fcb_tune(b"mac_impl", 0))


@pytest.mark.parametrize("B,L", [(64, 1000), (256, 3000)])
def test_ragged_call_sizes(F, B, L):
    """arbitrary call sizes / zero added latency (src/fft_convolver.rs:222-231)"""
    h = oracle.gen_ir(1, 0, L)
    x = oracle.gen_noise(1, 0, B * 40)
    rng = np.random.default_rng(11)
    sizes = [int(v) for v in rng.integers(1, 3 * B, size=64)]
    y = _run(F.FFTConvolver.init(h, B, L), x, sizes)
    yo = _run(oracle.FFTConvolver.init(h, B, L), x, sizes)
    yt = oracle_np.truth_f64(x, h)
    assert np.max(np.abs(y - yo)) <= TOL * rms(yt)
    assert np.max(np.abs(y - yt)) <= TOL * rms(yt)


def test_output_longer_input_allowed_and_empty_calls(F):
    B, L = 32, 100
    h = oracle.gen_ir(2, 0, L)
    g, o = F.FFTConvolver.init(h, B, L), oracle.FFTConvolver.init(h, B, L)
    x = oracle.gen_noise(2, 0, 200)
    og, oo = np.zeros(50, np.float32), np.zeros(50, np.float32)
    g.process(x[:66], og)  # input longer than output: only output.len() samples consumed (:222)
    o.process(x[:66], oo)
    assert np.max(np.abs(og - oo)) <= TOL * rms(oo)
    g.process(np.zeros(0, np.float32), np.zeros(0, np.float32))
    assert g.fill == o.fill == 50 % B


def test_update_changes_segment_count(F):
    """update() with a shorter IR: ring re-read modulo the new count (quirk 5), mid-block update"""
    B, L = 32, 320
    h0, h1, h2 = oracle.gen_ir(5, 0, L), oracle.gen_ir(5, 1, 100), oracle.gen_ir(5, 2, 300)
    x = oracle.gen_noise(5, 0, B * 60)
    g, o = F.FFTConvolver.init(h0, B, L), oracle.FFTConvolver.init(h0, B, L)
    og, oo = np.zeros(B, np.float32), np.zeros(B, np.float32)
    run = WholeRun()
    for i in range(60):
        if i == 13:
            g.update(h1); o.update(h1)
            assert g.active_seg_count == o.active_seg_count == 4
        if i == 30:
            g.update(h2); o.update(h2)
        if i == 40:  # mid-block update: feed half a block first
            g.process(x[i * B:i * B + 16], og[:2]); o.process(x[i * B:i * B + 16], oo[:2])
            g.update(h0); o.update(h0)
            g.process(x[i * B + 16:(i + 1) * B], og[16:]); o.process(x[i * B + 16:(i + 1) * B], oo[16:])
        else:
            g.process(x[i * B:(i + 1) * B], og); o.process(x[i * B:(i + 1) * B], oo)
        assert g.current == o.current
        run.add(og, oo)
    run.check(TOL, "update with changing segment counts")


def test_empty_ir_and_zero_length(F):
    g = F.FFTConvolver.init(np.zeros(0, np.float32), 16, 0)
    out = np.ones(10, np.float32)
    g.process(np.ones(10, np.float32), out)
    assert np.all(out == 0)  # active_seg_count == 0 -> zero fill (:216-219)
    g2 = F.FFTConvolver.init(np.ones(8, np.float32), 4, 8)
    g2.update(np.zeros(0, np.float32))
    assert g2.active_seg_count == 0
    g2.process(np.ones(10, np.float32), out)
    assert np.all(out == 0)


def test_batched_channels_match_mono(F):
    C, B, L = 7, 128, 1000
    irs = np.stack([oracle.gen_ir(c, 0, L) for c in range(C)])
    x = np.stack([oracle.gen_noise(c, 0, B * 20) for c in range(C)])
    yb = _run(F.FFTConvolver.init(irs, B, L), x, [B, 37, 200])
    for c in range(C):
        ym = _run(F.FFTConvolver.init(irs[c], B, L), x[c], [B, 37, 200])
        assert np.array_equal(yb[c], ym)
        yo = _run(oracle.FFTConvolver.init(irs[c], B, L), x[c], [B, 37, 200])
        assert np.max(np.abs(yb[c] - yo)) <= TOL * rms(yo)


def test_pipelined_host_path_large_batch(F):
    """>= 1024 channels with whole-block host calls take the copy/compute pipeline over channel
    groups (fcb_engine_process_block_host); it must give the bits of the plain path."""
    from fft_convolution_b200 import _lib
    C, B, L = 1024 + 37, 64, 200
    irs = np.stack([oracle.gen_ir(c % 5, c // 5, L) for c in range(C)])
    x = np.stack([oracle.gen_noise(c, 0, B * 9) for c in range(C)])
    _lib.check(_lib.load().fcb_tune(b"pipe_group", 200))
    yb = _run(F.FFTConvolver.init(irs, B, L), x, [B, B, 40, 24, B])  # mixes pipelined and chunked calls
    _lib.check(_lib.load().fcb_tune(b"pipe_group", 512))
    for c in (0, 199, 200, 511, 1023, 1024, C - 1):
        ym = _run(F.FFTConvolver.init(irs[c], B, L), x[c], [B, B, 40, 24, B])
        assert np.array_equal(yb[c], ym), c
        yo = _run(oracle.FFTConvolver.init(irs[c], B, L), x[c], [B, B, 40, 24, B])
        assert np.max(np.abs(yb[c] - yo)) <= TOL * rms(yo)


def test_pipelined_host_path_three_kernel_engine(F):
    """>= 1024 channels at B = 1024 (beyond the fused kernel): groups run K1, K2, K3 on the pipeline streams"""
    C, B, L = 1030, 1024, 2100
    h = oracle.gen_ir(8, 0, L)
    rng = np.random.default_rng(9)
    x = rng.uniform(-1, 1, size=(C, B * 4)).astype(np.float32)
    y = _run(F.FFTConvolver.init(h, B, L, channels=C), x, [B])
    for c in (0, 511, 512, C - 1):
        yo = _run(oracle.FFTConvolver.init(h, B, L), x[c], [B])
        assert np.max(np.abs(y[c] - yo)) <= TOL * rms(yo), c


@pytest.mark.parametrize("B,C", [(32, 19), (64, 9), (128, 5), (256, 3), (512, 2), (512, 1)])
def test_fused_block_kernel_is_bit_identical_to_three_kernels(F, B, C):
    """whole blocks with B in 32..512 run as ONE fused K1+K2+K3 kernel; it must give the bits of
    the K1 -> K2 -> K3 sequence (channel counts that leave a CTA partly empty included)"""
    from fft_convolution_b200 import _lib
    lib = _lib.load()
    L = B * 7 + 3
    irs = np.stack([oracle.gen_ir(c, 0, L) for c in range(C)])
    x = np.stack([oracle.gen_noise(c, 0, B * 11) for c in range(C)])
    outs = []
    for fused in (1, 0):
        _lib.check(lib.fcb_tune(b"fused_block", fused))
        conv = F.FFTConvolver.init(irs, B, L)
        y = _run(conv, x, [B, B, B // 2, B // 2, B])  # whole blocks and split blocks interleaved
        outs.append(y)
        if fused:  # ring / overlap state left behind must match too
            st_f = (conv.segment(conv.current, chan=C - 1).copy(), conv.overlap(chan=C - 1).copy())
        else:
            st_u = (conv.segment(conv.current, chan=C - 1).copy(), conv.overlap(chan=C - 1).copy())
    _lib.check(lib.fcb_tune(b"fused_block", 1))
    assert np.array_equal(outs[0], outs[1])
    assert np.array_equal(st_f[0], st_u[0]) and np.array_equal(st_f[1], st_u[1])
    yo = _run(oracle.FFTConvolver.init(irs[0], B, L), x[0], [B])
    assert np.max(np.abs(outs[0][0] - yo)) <= TOL * rms(yo)


def test_shared_ir_matches_per_channel_ir(F):
    C, B, L = 5, 64, 700
    h = oracle.gen_ir(9, 0, L)
    x = np.stack([oracle.gen_noise(c, 0, B * 16) for c in range(C)])
    ys = _run(F.FFTConvolver.init(h, B, L, channels=C), x, [B])
    yp = _run(F.FFTConvolver.init(np.tile(h, (C, 1)), B, L), x, [B])
    assert np.array_equal(ys, yp)


@pytest.mark.parametrize("B,C", [(128, 13), (256, 7), (512, 5), (512, 2)])
def test_shared_ir_reuse_kernel_bit_identical(F, B, C):
    """shared-IR engines with B >= 128 stage every IR tile once per CTA (k_block_fused_shared);
    same bits as one private copy of the IR per channel, and as the same engine with reuse off"""
    from fft_convolution_b200 import _lib
    lib = _lib.load()
    L = B * 9 + 11
    h = oracle.gen_ir(21, 0, L)
    x = np.stack([oracle.gen_noise(c, 0, B * 12) for c in range(C)])
    sizes = [B, B, B // 4, 3 * B // 4, B]
    ys = _run(F.FFTConvolver.init(h, B, L, channels=C), x, sizes)
    yp = _run(F.FFTConvolver.init(np.tile(h, (C, 1)), B, L), x, sizes)
    _lib.check(lib.fcb_tune(b"shared_reuse", 0))
    yn = _run(F.FFTConvolver.init(h, B, L, channels=C), x, sizes)
    _lib.check(lib.fcb_tune(b"shared_reuse", 1))
    assert np.array_equal(ys, yp) and np.array_equal(ys, yn)
    yo = _run(oracle.FFTConvolver.init(h, B, L), x[C - 1], sizes)
    assert np.max(np.abs(ys[C - 1] - yo)) <= TOL * rms(yo)


def test_zero_copy_pinned_buffers_match_copy_path(F):
    """>= 1024 channels with PINNED caller buffers: the kernel reads/writes host memory directly;
    same bits as the staged copy pipeline"""
    import ctypes as C
    from fft_convolution_b200 import _lib
    lib = _lib.load()
    Cn, B, L = 1100, 64, 300
    h = oracle.gen_ir(4, 0, L)
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, size=(Cn, B * 5)).astype(np.float32)
    p_in, p_out = lib.fcb_host_alloc(Cn * B * 4), lib.fcb_host_alloc(Cn * B * 4)
    h_in = np.ctypeslib.as_array(C.cast(p_in, C.POINTER(C.c_float)), shape=(Cn, B))
    h_out = np.ctypeslib.as_array(C.cast(p_out, C.POINTER(C.c_float)), shape=(Cn, B))
    ys = []
    for on in (1, 0):
        _lib.check(lib.fcb_tune(b"zero_copy", on))
        conv = F.FFTConvolver.init(h, B, L, channels=Cn)
        y = np.zeros_like(x)
        for b in range(5):
            h_in[:] = x[:, b * B:(b + 1) * B]
            conv.process(h_in, h_out)
            y[:, b * B:(b + 1) * B] = h_out
        ys.append(y)
    _lib.check(lib.fcb_tune(b"zero_copy", 1))
    lib.fcb_host_free(p_in)
    lib.fcb_host_free(p_out)
    assert np.array_equal(ys[0], ys[1])
    yo = _run(oracle.FFTConvolver.init(h, B, L), x[777], [B])
    assert np.max(np.abs(ys[0][777] - yo)) <= TOL * rms(yo)


def test_mapped_io_path_matches_copy_path(F):
    """small-batch host calls stage through mapped pinned memory; same bits as the copy-engine path"""
    from fft_convolution_b200 import _lib
    lib = _lib.load()
    C, B, L = 3, 128, 900
    irs = np.stack([oracle.gen_ir(c, 0, L) for c in range(C)])
    x = np.stack([oracle.gen_noise(c, 0, B * 10) for c in range(C)])
    sizes = [B, 50, 78, B, 200, 3]
    ys = []
    for on in (1, 0):
        _lib.check(lib.fcb_tune(b"mapped_io", on))
        ys.append(_run(F.FFTConvolver.init(irs, B, L), x, sizes))
    _lib.check(lib.fcb_tune(b"mapped_io", 1))
    assert np.array_equal(ys[0], ys[1])


def test_more_than_65535_channels(F):
    """grid and pointer arithmetic beyond 16-bit channel counts"""
    C, B, L = 70001, 32, 70
    h = oracle.gen_ir(3, 0, L)
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, size=(C, B * 3)).astype(np.float32)
    y = _run(F.FFTConvolver.init(h, B, L, channels=C), x, [B])
    for c in (0, 65535, 65536, C - 1):
        yo = _run(oracle.FFTConvolver.init(h, B, L), x[c], [B])
        assert np.max(np.abs(y[c] - yo)) <= TOL * max(rms(yo), 1e-3), c


def test_clone_is_deep(F):
    B, L = 64, 500
    h = oracle.gen_ir(3, 0, L)
    x = oracle.gen_noise(3, 0, B * 12)
    a = F.FFTConvolver.init(h, B, L)
    o1 = _run(a, x[:B * 6], [B])
    b = a.clone()
    ya = _run(a, x[B * 6:], [B])
    yb = _run(b, x[B * 6:], [B])
    assert np.array_equal(ya, yb)
    ref = _run(oracle.FFTConvolver.init(h, B, L), x, [B])
    assert np.max(np.abs(np.concatenate([o1, ya]) - ref)) <= TOL * rms(ref)


@pytest.mark.parametrize("async_tail", [False, True])
@pytest.mark.parametrize("H,L,sizes", [(64, 12000, [64]), (64, 12000, [17, 64, 5, 33]), (16, 5000, [16, 11, 3]),
                                      (128, 2000, [128]), (32, 40, [32])])
def test_twostage_vs_oracle(F, H, L, sizes, async_tail):
    """two-stage bookkeeping incl. ragged calls, non-power-of-two head, absent tail stages"""
    C = 3
    irs = np.stack([oracle.gen_ir(c, 0, L) for c in range(C)])
    x = np.stack([oracle.gen_noise(c, 0, 64 * 120) for c in range(C)])
    g = F.TwoStageFFTConvolver.init(irs, H, L, async_tail=async_tail)
    y = _run(g, x, sizes)
    for c in range(C):
        o = oracle.TwoStageFFTConvolver.init(irs[c], H, L)
        assert o.tail_block_size == g.tail_block_size
        yo = _run(o, x[c], sizes)
        yt = oracle_np.truth_f64(x[c], irs[c])
        assert np.max(np.abs(y[c] - yo)) <= TOL * rms(yt)
        assert np.max(np.abs(y[c] - yt)) <= TOL * rms(yt)


def test_twostage_non_power_of_two_head_panics_like_reference(F):
    """head 48 does not divide T = 512: the reference's tail_input slice (src/fft_convolver.rs:459)
    panics on the 11th block; engine and oracle must fail at the same call, agreeing until then."""
    H, L = 48, 5000
    h = oracle.gen_ir(0, 0, L)
    x = oracle.gen_noise(0, 0, H * 12)
    g, o = F.TwoStageFFTConvolver.init(h, H, L), oracle.TwoStageFFTConvolver.init(h, H, L)
    assert g.tail_block_size == o.tail_block_size == 512
    og, oo = np.zeros(H, np.float32), np.zeros(H, np.float32)
    failed_at = None
    run = WholeRun()
    for i in range(12):
        blk = x[i * H:(i + 1) * H]
        try:
            o.process(blk, oo)
        except oracle.OraclePanic:
            with pytest.raises(F.ConvolutionPanic):
                g.process(blk, og)
            failed_at = i
            break
        g.process(blk, og)
        run.add(og, oo)
    run.check(TOL, "two-stage, head size not dividing T")
    assert failed_at == 10


def test_twostage_config2_shape(F):
    """BASELINE configs[1] shape at reduced channel count: head 128, 5 s IR -> T = 8192 (derived),
    and the forced-T = 4096 variant, against the oracle with the same override."""
    H, L, C = 128, 240000, 2
    irs = np.stack([oracle.gen_ir(c, 0, L) for c in range(C)])
    x = np.stack([oracle.gen_noise(c, 0, H * 200) for c in range(C)])
    for forced in (0, 4096):
        g = F.TwoStageFFTConvolver.init(irs, H, L, forced_tail_block=forced, async_tail=True)
        assert g.tail_block_size == (forced or 8192)
        y = _run(g, x, [H])
        o = oracle.TwoStageFFTConvolver.init(irs[1], H, L, forced_tail=forced)
        yo = _run(o, x[1], [H])
        assert np.max(np.abs(y[1] - yo)) <= TOL * rms(yo)


def test_crossfade_sequences_vs_oracle(F):
    """crossfade state machine + gain law: fades, pending updates, short outputs; C = 2 channels"""
    B, L, C = 64, 300, 2
    irs = [np.stack([oracle.gen_ir(c, u, L) for c in range(C)]) for u in range(4)]
    x = np.stack([oracle.gen_noise(c, 0, B * 80) for c in range(C)])
    g = F.CrossfadeConvolver.new(F.FFTConvolver.init(irs[0], B, L), L, B, 200)
    os_ = [oracle.CrossfadeConvolver.new(oracle.FFTConvolver.init(irs[0][c], B, L), L, B, 200) for c in range(C)]
    og = np.zeros((C, B), np.float32)
    oo = np.zeros(B, np.float32)
    upd = {5: 1, 7: 2, 8: 3, 30: 1, 50: 2}
    runs = [WholeRun() for _ in range(C)]
    for i in range(80):
        if i in upd:
            g.update(irs[upd[i]])
            for c in range(C):
                os_[c].update(irs[upd[i]][c])
        n_out = B if i % 9 else B - 13  # sometimes a short output: crossfader advances by out.len() only
        blk = np.ascontiguousarray(x[:, i * B:(i + 1) * B])
        og = np.zeros((C, n_out), np.float32)
        g.process(blk, og)
        for c in range(C):
            oo[:] = 0
            os_[c].process(blk[c], oo[:n_out])
            runs[c].add(og[c], oo[:n_out])
            assert g.is_crossfading() == os_[c].is_crossfading()
            cnt, mix, appr, tgt = g.state()
            s = os_[c].crossfader
            assert (cnt, appr, tgt) == (s.counter, bool(s.approaching), s.target)
            assert np.float32(mix) == np.float32(s.mix_value)
    for c in range(C):
        runs[c].check(TOL, f"crossfade sequence, channel {c}")


def test_crossfade_config3_shape(F):
    """BASELINE configs[2] shape, reduced: CrossfadeConvolver::init(h, 512, 96 000) => 96 000-sample
    fade + 512 hold, update every 50 blocks; 2 channels, 120 blocks."""
    B, L, C = 512, 96000, 2
    ir = lambda u: np.stack([oracle.gen_ir(c, u, L) for c in range(C)])
    x = np.stack([oracle.gen_noise(c, 0, B * 120) for c in range(C)])
    g = F.CrossfadeConvolver.init(ir(0), B, L)
    o = oracle.CrossfadeConvolver.init(ir(0)[1], B, L)
    og, oo = np.zeros((C, B), np.float32), np.zeros(B, np.float32)
    run = WholeRun()
    for i in range(120):
        if i and i % 50 == 0:
            g.update(ir(i // 50)); o.update(ir(i // 50)[1])
        blk = np.ascontiguousarray(x[:, i * B:(i + 1) * B])
        g.process(blk, og); o.process(blk[1], oo)
        run.add(og[1], oo)
    run.check(TOL, "configs[2] shape")


def test_engine_on_second_device(F):
    """every entry point selects its own device (skipped on a 1-GPU box)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    B, L = 8192, 20000  # a block size that needs the >48 KB shared-memory opt-in on each device
    h = oracle.gen_ir(1, 0, L)
    x = oracle.gen_noise(1, 0, B * 4)
    y0 = _run(F.FFTConvolver.init(h, B, L, device=0), x, [B])
    y1 = _run(F.FFTConvolver.init(h, B, L, device=1), x, [B])
    assert np.array_equal(y0, y1)
    h2 = oracle.gen_ir(2, 0, 3000)
    x2 = oracle.gen_noise(2, 0, 512 * 6)
    assert np.array_equal(_run(F.FFTConvolver.init(h2, 512, 3000, device=1), x2, [512]),
                          _run(F.FFTConvolver.init(h2, 512, 3000, device=0), x2, [512]))
