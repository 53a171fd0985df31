"""Round-2 real-time surface: the two methods the reference leaves todo!() (as documented extensions, semantics
fixed in the oracle first), background IR updates (shadow spectra + pointer flip), CrossfadeConvolver's Clone, and
the rule that update() / process() never allocate (src/lib.rs:8)."""
import ctypes as C

import numpy as np
import pytest

import oracle
from refsignals import rms

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    import fft_convolution_b200 as f
    return f


def _pinned(F, arr):
    """a page-locked copy of `arr` (fcb_host_alloc) as (numpy view, raw pointer)"""
    lib = F.load_library()
    a = np.ascontiguousarray(arr, np.float32)
    p = lib.fcb_host_alloc(a.nbytes)
    v = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=a.shape)
    v[...] = a
    return v, p


@pytest.mark.parametrize("head,L,upd_len", [(64, 12000, 12000), (64, 12000, 5000), (32, 700, 700), (128, 300, 120)])
def test_twostage_update_extension_matches_oracle(F, head, L, upd_len):
    """TwoStageFFTConvolver::update (todo!() in the reference): per-stage update on the re-sliced response"""
    h0, h1 = oracle.gen_ir(3, 0, L), oracle.gen_ir(3, 1, upd_len)
    g, o = F.TwoStageFFTConvolver.init(h0, head, L), oracle.TwoStageFFTConvolver.init(h0, head, L)
    n = 90
    x = oracle.gen_noise(3, 0, head * n)
    yg, yo = np.zeros_like(x), np.zeros_like(x)
    a, b = np.zeros(head, np.float32), np.zeros(head, np.float32)
    for i in range(n):
        if i in (20, 57):
            g.update(h1 if i == 20 else h0)
            o.update_ext(h1 if i == 20 else h0)
        g.process(x[i * head:(i + 1) * head], a)
        o.process(x[i * head:(i + 1) * head], b)
        yg[i * head:(i + 1) * head], yo[i * head:(i + 1) * head] = a, b
    assert np.max(np.abs(yg - yo)) <= 1e-5 * rms(yo)


def test_twostage_update_extension_batched_equals_mono(F):
    C_, head, L = 3, 64, 9000
    h0 = np.stack([oracle.gen_ir(c, 0, L) for c in range(C_)])
    h1 = np.stack([oracle.gen_ir(c, 1, L) for c in range(C_)])
    x = np.stack([oracle.gen_noise(c, 0, head * 60) for c in range(C_)])
    g = F.TwoStageFFTConvolver.init(h0, head, L)
    y = np.zeros_like(x)
    blk = np.zeros((C_, head), np.float32)
    for i in range(60):
        if i == 25:
            g.update(h1)
        g.process(np.ascontiguousarray(x[:, i * head:(i + 1) * head]), blk)
        y[:, i * head:(i + 1) * head] = blk
    ref = oracle.batch_twostage(h0, head, x, head, irs_upd=h1[None], update_every=25)
    # the oracle batch applies the update every 25 calls (25, 50): replay that schedule for the second one as well
    g2 = F.TwoStageFFTConvolver.init(h0, head, L)
    y2 = np.zeros_like(x)
    for i in range(60):
        if i in (25, 50):
            g2.update(h1)
        g2.process(np.ascontiguousarray(x[:, i * head:(i + 1) * head]), blk)
        y2[:, i * head:(i + 1) * head] = blk
    for c in range(C_):
        assert np.max(np.abs(y2[c] - ref[c])) <= 1e-5 * rms(ref[c])
    assert np.array_equal(y[:, :head * 50], y2[:, :head * 50])


@pytest.mark.parametrize("reset_at", [3, 9, 12, 30])
def test_crossfade_reset_extension_matches_oracle(F, reset_at):
    """CrossfadeConvolver::reset (todo!() in the reference): idle, holding, fading and with a pending response"""
    B, L, fade = 64, 500, 300
    h = [oracle.gen_ir(5, u, L) for u in range(3)]
    g = F.CrossfadeConvolver.new(F.FFTConvolver.init(h[0], B, L), L, B, fade)
    o = oracle.CrossfadeConvolver.new(oracle.FFTConvolver.init(h[0], B, L), L, B, fade)
    a, b = np.zeros(B, np.float32), np.zeros(B, np.float32)
    worst, scale = 0.0, 0.05
    for i in range(40):
        if i == 8:
            g.update(h[1]); o.update(h[1])
        if i == 11:  # arrives while the first fade is running: stored, pending
            g.update(h[2]); o.update(h[2])
        if i == reset_at:
            g.reset(); o.reset_ext()
        x = oracle.gen_noise(5, i * B, B)
        g.process(x, a); o.process(x, b)
        scale = max(scale, rms(b))
        worst = max(worst, float(np.max(np.abs(a - b))))
        cnt, mix, appr, tgt = g.state()
        s = o.crossfader
        assert (cnt, appr, tgt, np.float32(mix)) == (s.counter, bool(s.approaching), s.target, np.float32(s.mix_value)), i
        assert g.is_crossfading() == o.is_crossfading()
    assert worst <= 1e-5 * scale


def test_crossfade_clone_is_a_deep_copy(F):
    B, L, fade = 64, 400, 200
    h = [oracle.gen_ir(6, u, L) for u in range(2)]
    g = F.CrossfadeConvolver.new(F.FFTConvolver.init(h[0], B, L), L, B, fade)
    a, b = np.zeros(B, np.float32), np.zeros(B, np.float32)
    for i in range(7):
        if i == 4:
            g.update(h[1])
        g.process(oracle.gen_noise(6, i * B, B), a)
    k = g.clone()  # mid-fade
    assert k.state() == g.state()
    for i in range(7, 20):
        x = oracle.gen_noise(6, i * B, B)
        g.process(x, a); k.process(x, b)
        assert np.array_equal(a, b), i
    k.process(oracle.gen_noise(9, 0, B), b)  # the clone moves on alone; the original is untouched
    x = oracle.gen_noise(6, 20 * B, B)
    g.process(x, a)
    o = oracle.CrossfadeConvolver.new(oracle.FFTConvolver.init(h[0], B, L), L, B, fade)
    for i in range(21):
        if i == 4:
            o.update(h[1])
        o.process(oracle.gen_noise(6, i * B, B), b)
    assert np.max(np.abs(a - b)) <= 1e-5 * max(rms(b), 0.05)


def test_shared_ir_crossfade_update_while_fading(F):
    """a crossfade built over a shared-IR convolver is handed ONE response row, also when update() arrives mid-fade
    (round-1 advisor finding: the stored-response copy read C rows)"""
    C_, B, L, fade = 5, 64, 600, 400
    h = [oracle.gen_ir(8, u, L) for u in range(3)]
    g = F.CrossfadeConvolver.new(F.FFTConvolver.init(h[0], B, L, channels=C_), L, B, fade)
    os_ = [oracle.CrossfadeConvolver.new(oracle.FFTConvolver.init(h[0], B, L), L, B, fade) for _ in range(C_)]
    blk, ref = np.zeros((C_, B), np.float32), np.zeros(B, np.float32)
    for i in range(30):
        if i == 3:
            g.update(h[1]); [o.update(h[1]) for o in os_]
        if i == 5:  # mid-fade: goes to stored_response
            g.update(h[2][:L - 7]); [o.update(h[2][:L - 7]) for o in os_]
        x = np.stack([oracle.gen_noise(30 + c, i * B, B) for c in range(C_)])
        g.process(x, blk)
        for c in range(C_):
            os_[c].process(x[c], ref)
            assert np.max(np.abs(blk[c] - ref)) <= 1e-5 * max(rms(ref), 0.05), (i, c)


def test_background_update_in_wait_mode_equals_update(F):
    """fcb_fftconv_update_begin(FCB_UPDATE_WAIT): committed at the very next process call — bit-identical to update()"""
    C_, B, L = 4, 128, 128 * 9 + 17
    h0 = np.stack([oracle.gen_ir(c, 0, L) for c in range(C_)])
    h1 = np.stack([oracle.gen_ir(c, 1, L - 300) for c in range(C_)])
    x = np.stack([oracle.gen_noise(c, 0, B * 30) for c in range(C_)])
    hv, hp = _pinned(F, h1)
    ga, gb = F.FFTConvolver.init(h0, B, L), F.FFTConvolver.init(h0, B, L)
    gb.update_reserve()
    ya, yb = np.zeros((C_, B), np.float32), np.zeros((C_, B), np.float32)
    for i in range(30):
        if i in (7, 19):
            ga.update(h1)
            gb.update_begin(hp, h1.shape[1], wait=True)
            assert gb.update_pending()
        blkx = np.ascontiguousarray(x[:, i * B:(i + 1) * B])
        ga.process(blkx, ya); gb.process(blkx, yb)
        assert not gb.update_pending()
        assert np.array_equal(ya, yb), i
    assert ga.active_seg_count == gb.active_seg_count
    F.load_library().fcb_host_free(hp)


def test_background_update_flips_between_two_blocks(F):
    """without FCB_UPDATE_WAIT the new response is swapped in by the first process call that finds K5 finished; whatever
    that call is, the output equals the oracle's with update() placed before the same call"""
    B, L = 256, 256 * 40
    h0, h1 = oracle.gen_ir(1, 0, L), oracle.gen_ir(1, 1, L)
    hv, hp = _pinned(F, h1)
    g, o = F.FFTConvolver.init(h0, B, L), oracle.FFTConvolver.init(h0, B, L)
    g.update_reserve()
    a, b = np.zeros(B, np.float32), np.zeros(B, np.float32)
    flipped_at = None
    for i in range(60):
        if i == 10:
            g.update_begin(hp, L)
        was_pending = g.update_pending()
        x = oracle.gen_noise(1, i * B, B)
        g.process(x, a)
        if was_pending and not g.update_pending():
            flipped_at = i
            o.update(h1)
        o.process(x, b)
        assert np.max(np.abs(a - b)) <= 1e-5 * max(rms(b), 0.05), i
    assert flipped_at is not None and flipped_at >= 10
    F.load_library().fcb_host_free(hp)


def test_update_and_process_do_not_allocate_in_the_steady_state(F):
    """src/lib.rs:8 — `update` must be real-time safe (no allocations); so must process().  The library counts every
    cudaMalloc / cudaHostAlloc / stream / event creation and release it makes (fcb_debug_alloc_count)."""
    B, L, C_ = 64, 64 * 12, 3
    h = np.stack([oracle.gen_ir(c, 0, L) for c in range(C_)])
    x = np.stack([oracle.gen_noise(c, 0, B * 8) for c in range(C_)])
    hv, hp = _pinned(F, h)
    uni = F.FFTConvolver.init(h, B, L)
    uni.update_reserve()
    uni.reserve(B * 8)
    two = F.TwoStageFFTConvolver.init(h, 32, L, async_tail=True)
    xf = F.CrossfadeConvolver.init(h, B, L)
    big = F.FFTConvolver.init(np.stack([oracle.gen_ir(0, 0, 96)] * 1100), 32, 96)  # >= 1024 channels: grouped host pipeline
    xbig = np.zeros((1100, 32), np.float32)
    ybig = np.zeros((1100, 32), np.float32)
    out1, out8 = np.zeros((C_, B), np.float32), np.zeros((C_, B * 8), np.float32)
    out32, outr = np.zeros((C_, 32), np.float32), np.zeros((C_, 23), np.float32)

    def cycle(k):
        uni.process(np.ascontiguousarray(x[:, :B]), out1)           # whole block (fused kernel)
        uni.process(np.ascontiguousarray(x[:, :23]), outr)          # ragged chunk (three kernels)
        uni.process(np.ascontiguousarray(x[:, :B - 23]), np.zeros((C_, B - 23), np.float32))
        uni.process(x, out8)                                        # eight blocks in one call (time-batched pass)
        uni.update(h)
        uni.update_begin(hp, L, wait=bool(k & 1))
        uni.process(np.ascontiguousarray(x[:, :B]), out1)
        uni.reset()
        for j in range(3):
            two.process(np.ascontiguousarray(x[:, j * 32:(j + 1) * 32]), out32)
        two.update(h)
        two.reset()
        xf.process(np.ascontiguousarray(x[:, :B]), out1)
        xf.update(h)                                                # swap, or stored while fading
        xf.process(np.ascontiguousarray(x[:, B:2 * B]), out1)
        xf.reset()
        big.process(xbig, ybig)

    for k in range(3):  # first uses may still create lazily initialised objects of the CUDA runtime, not ours
        cycle(k)
    before = F.alloc_count()
    for k in range(6):
        cycle(k)
    assert F.alloc_count() == before
    F.load_library().fcb_host_free(hp)


@pytest.mark.parametrize("kind,C_,B,L", [("uniform", 1, 256, 48000), ("uniform", 3, 512, 20000), ("uniform", 20, 64, 3000),
                                          ("twostage", 8, 128, 40000), ("crossfade", 6, 512, 30000)])
def test_small_batch_split_is_deterministic_and_within_tolerance(F, kind, C_, B, L):
    """small batches cut each delay line over several CTAs (partials summed in slice order by the last CTA to arrive):
    two runs give the same bits, and the result stays within 1e-5 x RMS of the oracle"""
    from refsignals import WholeRun
    h = np.stack([oracle.gen_ir(c, 0, L) for c in range(C_)])
    h1 = np.stack([oracle.gen_ir(c, 1, L) for c in range(C_)])
    nblk = 40
    x = np.stack([oracle.gen_noise(c, 0, B * nblk) for c in range(C_)])

    def run_gpu():
        if kind == "uniform":
            g = F.FFTConvolver.init(h, B, L)
        elif kind == "twostage":
            g = F.TwoStageFFTConvolver.init(h, B, L)
        else:
            g = F.CrossfadeConvolver.new(F.FFTConvolver.init(h, B, L), L, B, 3 * B)
        y = np.zeros_like(x)
        blk = np.zeros((C_, B), np.float32)
        for i in range(nblk):
            if kind == "crossfade" and i == 9:
                g.update(h1)
            g.process(np.ascontiguousarray(x[:, i * B:(i + 1) * B]), blk)
            y[:, i * B:(i + 1) * B] = blk
        return y

    y1, y2 = run_gpu(), run_gpu()
    assert np.array_equal(y1, y2)
    for c in range(C_):
        if kind == "uniform":
            o = oracle.FFTConvolver.init(h[c], B, L)
        elif kind == "twostage":
            o = oracle.TwoStageFFTConvolver.init(h[c], B, L)
        else:
            o = oracle.CrossfadeConvolver.new(oracle.FFTConvolver.init(h[c], B, L), L, B, 3 * B)
        run = WholeRun()
        blk = np.zeros(B, np.float32)
        for i in range(nblk):
            if kind == "crossfade" and i == 9:
                o.update(h1[c])
            o.process(x[c, i * B:(i + 1) * B], blk)
            run.add(y1[c, i * B:(i + 1) * B], blk)
        run.check(1e-5, f"{kind} channel {c}")


def test_small_batch_calls_on_caller_pinned_buffers_work_in_place(F):
    """page-locked caller buffers (fcb_host_alloc) are read and written by the kernels in place — same result as the
    staged path, for all three convolver types"""
    C_, B, L = 6, 128, 5000
    h = np.stack([oracle.gen_ir(c, 0, L) for c in range(C_)])
    lib = F.load_library()
    xin, pin = _pinned(F, np.zeros((C_, B), np.float32))
    yout, pout = _pinned(F, np.zeros((C_, B), np.float32))
    for make in (lambda: F.FFTConvolver.init(h, B, L), lambda: F.TwoStageFFTConvolver.init(h, B, L),
                 lambda: F.CrossfadeConvolver.new(F.FFTConvolver.init(h, B, L), L, B, 3 * B)):
        a, b = make(), make()
        ref = np.zeros((C_, B), np.float32)
        for i in range(12):
            x = np.stack([oracle.gen_noise(c, i * B, B) for c in range(C_)])
            if i == 5 and hasattr(a, "is_crossfading"):
                h1 = np.stack([oracle.gen_ir(c, 1, L) for c in range(C_)])
                a.update(h1); b.update(h1)
            xin[...] = x
            from fft_convolution_b200 import _lib
            fn = {"FFTConvolver": lib.fcb_fftconv_process, "TwoStageFFTConvolver": lib.fcb_twostage_process,
                  "CrossfadeConvolver": lib.fcb_crossfade_process}[type(a).__name__]
            _lib.check(fn(a._h, pin, B, B, pout, B, B))
            b.process(x, ref)
            assert np.array_equal(np.asarray(yout), ref), (type(a).__name__, i)
    lib.fcb_host_free(pin)
    lib.fcb_host_free(pout)


@pytest.mark.parametrize("head,L,stages,async_tail", [(32, 40000, 3, False), (32, 40000, 4, True), (64, 150000, 3, True)])
def test_nested_partition_matches_oracle_and_truth(F, head, L, stages, async_tail):
    """fcb_options.stages > 2 (extension): the tail of the two-stage partition is again a two-stage
    convolver — block sizes head, T1, T2, ... — against the oracle's nested convolver and an f64 convolution"""
    from oracle import oracle_np
    from refsignals import WholeRun
    h = oracle.gen_ir(2, 0, L)
    g = F.TwoStageFFTConvolver.init(h, head, L, stages=stages, async_tail=async_tail)
    o = oracle.TwoStageFFTConvolver.init(h, head, L, stages=stages, max_block=16384)
    assert g.stage_blocks == o.stage_blocks and len(g.stage_blocks) >= 3 and g.stage_blocks[2] > g.stage_blocks[1] >= head
    n = 3 * g.stage_blocks[2] // head + 40
    x = oracle.gen_noise(2, 0, head * n)
    yg, yo = np.zeros_like(x), np.zeros_like(x)
    a, b = np.zeros(head, np.float32), np.zeros(head, np.float32)
    for i in range(n):
        g.process(x[i * head:(i + 1) * head], a)
        o.process(x[i * head:(i + 1) * head], b)
        yg[i * head:(i + 1) * head], yo[i * head:(i + 1) * head] = a, b
    assert np.max(np.abs(yg - yo)) <= 1e-5 * rms(yo)
    yt = oracle_np.truth_f64(x, h)
    assert np.max(np.abs(yg - yt)) <= 1e-5 * rms(yt)
    # reset and clone carry the nested state
    k = g.clone()
    g.process(x[:head], a); k.process(x[:head], b)
    assert np.array_equal(a, b)
    g.reset(); o.reset()
    for i in range(20):
        g.process(x[i * head:(i + 1) * head], a); o.process(x[i * head:(i + 1) * head], b)
    assert np.max(np.abs(a - b)) <= 1e-5 * rms(yo)
