"""The synchronous crossfade host call computes the NEXT block's gains while the GPU works on the current block
(fcb_crossfade::spec_gains, host_mirror.cu).  That must be invisible: same output bits, same crossfader state after every
call as with the gains computed on the critical path (fcb_tune("xf_speculate", 0)), through fades, updates that land
mid-fade (pending response, src/crossfade_convolver.rs:58-70), short outputs (the crossfader advances by output.len()
only, :75), reset (extension) and clone."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    import fft_convolution_b200 as f
    return f


def _tune(key, value):
    from fft_convolution_b200 import _lib
    _lib.check(_lib.load().fcb_tune(key, value))


def _sequence(F, speculate, C, pinned=False):
    B, L, fade = 64, 300, 200
    _tune(b"xf_speculate", 1 if speculate else 0)
    try:
        irs = [np.stack([oracle.gen_ir(c, u, L) for c in range(C)]) for u in range(4)]
        x = np.stack([oracle.gen_noise(c, 0, B * 90) for c in range(C)])
        g = F.CrossfadeConvolver.new(F.FFTConvolver.init(irs[0] if C > 1 else irs[0][0], B, L), L, B, fade)
        upd = {5: 1, 7: 2, 8: 3, 30: 1, 50: 2, 70: 3}
        outs, states = [], []
        for i in range(90):
            if i in upd:
                g.update(irs[upd[i]] if C > 1 else irs[upd[i]][0])
            if i == 60:
                g.reset()
            if i == 75:
                g = g.clone()
            n_out = B if i % 9 else B - 13
            blk = np.ascontiguousarray(x[:, i * B:(i + 1) * B])
            og = np.zeros((C, n_out), np.float32)
            g.process(blk if C > 1 else blk[0], og if C > 1 else og[0])
            outs.append(og.copy())
            states.append(g.state() + (g.is_crossfading(),))
        return outs, states
    finally:
        _tune(b"xf_speculate", 1)


@pytest.mark.parametrize("C", [1, 3])
def test_speculated_gains_are_invisible(F, C):
    a_out, a_st = _sequence(F, True, C)
    b_out, b_st = _sequence(F, False, C)
    assert a_st == b_st
    for i, (a, b) in enumerate(zip(a_out, b_out)):
        assert np.array_equal(a, b), f"block {i}"
    assert any(s[2] for s in a_st) and any(not s[2] for s in a_st)  # the run saw fades and rests


def test_speculated_gains_vs_oracle_state(F):
    """whole 512-sample blocks through a long fade (the BASELINE configs[2] pattern: update every 50 blocks, fade as long
    as the response), state compared with the CPU restatement after every call"""
    B, L, C = 128, 128 * 6, 2
    irs = [np.stack([oracle.gen_ir(c, u, L) for c in range(C)]) for u in range(3)]
    g = F.CrossfadeConvolver.init(irs[0], B, L)
    o = oracle.CrossfadeConvolver.init(irs[0][0], B, L)
    og, oo = np.zeros((C, B), np.float32), np.zeros(B, np.float32)
    err = ref = 0.0
    for i in range(40):
        if i % 10 == 9:
            g.update(irs[(i // 10 + 1) % 3])
            o.update(irs[(i // 10 + 1) % 3][0])
        blk = np.stack([oracle.gen_noise(c, i * B, B) for c in range(C)])
        g.process(blk, og)
        o.process(blk[0], oo)
        err = max(err, float(np.max(np.abs(og[0] - oo))))
        ref += float(np.sum(oo.astype(np.float64) ** 2))
        cnt, mix, appr, tgt = g.state()
        s = o.crossfader
        assert (cnt, appr, tgt, np.float32(mix)) == (s.counter, bool(s.approaching), s.target, np.float32(s.mix_value))
    assert err <= 1e-5 * np.sqrt(ref / (40 * B))
