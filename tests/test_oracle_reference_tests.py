"""Every test the reference holds for the hot path, restated against the CPU oracle.

This is what pins the oracle: the three delta-IR known answers,
test_crossfader's exact equalities and the six behavioural tests of src/tests.rs, with the
reference's own tolerances.  CPU only.
"""
import numpy as np
import pytest

import oracle
from oracle import oracle_np
from refsignals import generate_sinusoid

IMPLS = {
    "c": (oracle.FFTConvolver, oracle.TwoStageFFTConvolver),
    "np": (oracle_np.FFTConvolverNP, oracle_np.TwoStageNP),
}


@pytest.fixture(params=["c", "np"])
def impl(request):
    return IMPLS[request.param]


def _delta(n=1024):
    r = np.zeros(n, np.float32)
    r[0] = 1.0
    return r


def test_fft_convolver_passthrough(impl):
    """src/fft_convolver.rs:309-321"""
    conv = impl[0].init(_delta(), 1024, 1024)
    out = np.zeros(1024, np.float32)
    conv.process(np.ones(1024, np.float32), out)
    assert np.all(np.abs(out - 1.0) < 1e-6)


def test_fft_twostage_convolver_passthrough(impl):
    """src/fft_convolver.rs:528-540"""
    conv = impl[1].init(_delta(), 1024, 1024)
    out = np.zeros(1024, np.float32)
    conv.process(np.ones(1024, np.float32), out)
    assert np.all(np.abs(out - 1.0) < 1e-6)


def test_crossfade_convolver_passthrough():
    """src/crossfade_convolver.rs:107-124"""
    conv = oracle.CrossfadeConvolver.new(oracle.FFTConvolver.init(_delta(), 1024, 1024), 1024, 1024, 1024)
    out = np.zeros(1024, np.float32)
    conv.process(np.ones(1024, np.float32), out)
    assert np.all(np.abs(out - 1.0) < 1e-6)


def test_crossfader():
    """src/crossfade_convolver.rs:281-316 — exact equalities"""
    hold, fading = 4, 4
    a, b = np.float32(1.0), np.float32(10.0)
    x = oracle.Crossfader(fading, hold)
    start = {x.A: b, x.B: a}
    end = {x.A: a, x.B: b}
    for target in (x.B, x.A):
        x.fade_into(target)
        for i in range(hold + fading):
            v = x.mix(a, b)
            if i < hold:
                assert x.state == ("Approaching", target)
                assert v == start[target]
            elif i < hold + fading - 1:
                assert x.state == ("Approaching", target)
                assert v != start[target] and v != end[target]
            else:
                assert v == end[target]
                assert x.state == ("Reached", target)


def test_fft_convolver_update_is_reset(impl):
    """src/tests.rs:18-59"""
    F = impl[0]
    bs = 512
    ra = generate_sinusoid(bs, 1000.0, gain=1.0)
    rb = generate_sinusoid(bs, 2000.0, gain=0.7)
    ca, cb, cu = F.init(ra, bs, bs), F.init(rb, bs, bs), F.init(ra, bs, bs)
    oa, ob, ou = (np.zeros(bs, np.float32) for _ in range(3))
    x = generate_sinusoid(16 * bs, 1300.0)
    for i in range(16):
        if i == 8:
            cu.update(rb)
        blk = x[i * bs:(i + 1) * bs]
        cu.process(blk, ou)
        if i < 8:
            ca.process(blk, oa)
            assert np.all(np.abs(oa - ou) < 1e-6)
        else:
            cb.process(blk, ob)
            assert np.all(np.abs(ob - ou) < 1e-6)


def test_crossfade_convolver():
    """src/tests.rs:61-117"""
    bs = 512
    ra = generate_sinusoid(bs, 1000.0, gain=1.0)
    rb = generate_sinusoid(bs, 2000.0, gain=0.7)
    ca = oracle.FFTConvolver.init(ra, bs, bs)
    cb = oracle.FFTConvolver.init(rb, bs, bs)
    xf = oracle.CrossfadeConvolver.new(ca.clone(), bs, bs, bs)
    oa, ob, ox = (np.zeros(bs, np.float32) for _ in range(3))
    x = generate_sinusoid(16 * bs, 1300.0)
    for i in range(16):
        if i == 8:
            xf.update(rb)
        blk = x[i * bs:(i + 1) * bs]
        xf.process(blk, ox)
        ca.process(blk, oa)
        if i >= 8:
            cb.process(blk, ob)
        if i <= 8:
            assert np.all(np.abs(oa - ox) < 1e-6)
        elif i == 9:
            k = bs // 2 - 1
            assert abs(ox[k] - (oa[k] * np.float32(0.5) + ob[k] * np.float32(0.5))) < 1e-6
        else:
            assert np.all(np.abs(ob - ox) < 1e-6)


def test_block_size_equal(impl):
    """src/tests.rs:119-146"""
    F = impl[0]
    bs, nblocks = 128, 1000
    r = generate_sinusoid(bs, 1000.0, gain=0.1)
    ca, cb = F.init(r, bs // 2, bs), F.init(r, bs, bs)
    oa, ob = np.zeros(bs, np.float32), np.zeros(bs, np.float32)
    x = generate_sinusoid(nblocks * bs, 1300.0, gain=0.1)
    for i in range(nblocks):
        ca.process(x[i * bs:(i + 1) * bs], oa)
        cb.process(x[i * bs:(i + 1) * bs], ob)
        assert np.all(np.abs(oa - ob) < 1e-5)


def test_twostage_equal(impl):
    """src/tests.rs:148-175"""
    F, T = impl
    bs, nblocks = 64, 1000
    r = generate_sinusoid(12000, 1000.0, gain=0.1)
    ca, cb = F.init(r, bs // 2, r.size), T.init(r, bs, r.size)
    assert cb.tail_block_size == 1024
    oa, ob = np.zeros(bs, np.float32), np.zeros(bs, np.float32)
    x = generate_sinusoid(nblocks * bs, 1300.0, gain=0.1)
    for i in range(nblocks):
        ca.process(x[i * bs:(i + 1) * bs], oa)
        cb.process(x[i * bs:(i + 1) * bs], ob)
        assert np.all(np.abs(oa - ob) < 1e-5)


@pytest.mark.parametrize("kind", ["uniform", "twostage"])
def test_reset(kind):
    """src/tests.rs:177-216 and :218-257 (the reference compares only the first block; we
    compare the whole run)"""
    bs, nblocks = 64, 1000
    r = generate_sinusoid(12000, 1000.0, gain=0.1)
    cls = oracle.FFTConvolver if kind == "uniform" else oracle.TwoStageFFTConvolver
    conv = cls.init(r, bs, r.size)
    x = generate_sinusoid(nblocks * bs, 1300.0, gain=0.1)
    outs = []
    for _ in range(2):
        o = np.zeros(nblocks * bs, np.float32)
        blk = np.zeros(bs, np.float32)
        for i in range(nblocks):
            conv.process(x[i * bs:(i + 1) * bs], blk)
            o[i * bs:(i + 1) * bs] = blk
        outs.append(o)
        conv.reset()
    assert np.all(np.abs(outs[0] - outs[1]) < 1e-5)


def test_reference_panics():
    """contract violations that panic in the reference (src/fft_convolver.rs:106-110,
    191-193, 422-424, 428; src/crossfade_convolver.rs:80-82)"""
    with pytest.raises(oracle.OraclePanic):
        oracle.FFTConvolver.init(np.zeros(10, np.float32), 4, 5)
    c = oracle.FFTConvolver.init(np.zeros(10, np.float32), 4, 10)
    with pytest.raises(oracle.OraclePanic):
        c.update(np.zeros(11, np.float32))
    t = oracle.TwoStageFFTConvolver.init(np.zeros(100, np.float32), 8, 100)
    with pytest.raises(oracle.OraclePanic):
        t.update(np.zeros(10, np.float32))
    with pytest.raises(oracle.OraclePanic):
        t.process(np.zeros(9, np.float32), np.zeros(9, np.float32))
    x = oracle.CrossfadeConvolver.init(np.ones(16, np.float32), 8, 16)
    with pytest.raises(oracle.OraclePanic):
        x.reset()
