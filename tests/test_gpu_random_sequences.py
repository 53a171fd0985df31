"""Differential testing: seeded random sequences of process / update / reset / clone calls with
random call sizes, CUDA engine vs CPU oracle — outputs within 1e-5 x the whole run's output RMS and the host scheduler
scalars (current, fill, active segment count, crossfader state) identical at every step."""
import numpy as np
import pytest

import oracle
from refsignals import WholeRun

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    import fft_convolution_b200 as f
    return f


TOL = 1e-5  # north_star: max-abs error relative to the output RMS, taken over the whole run


@pytest.mark.parametrize("seed", range(12))
def test_fftconvolver_random_ops(F, seed):
    rng = np.random.default_rng(1000 + seed)
    B = int(2 ** rng.integers(3, 10))
    L = int(rng.integers(1, 12 * B))
    g = F.FFTConvolver.init(oracle.gen_ir(seed, 0, L), int(B - rng.integers(0, B // 2)), L)
    o = oracle.FFTConvolver.init(oracle.gen_ir(seed, 0, L), B, L)
    assert g.block_size == o.block_size == B
    pos, upd = 0, 1
    run = WholeRun()
    for step in range(70):
        op = rng.choice(["process"] * 8 + ["update", "reset", "clone"])
        if op == "process":
            n = int(rng.choice([B, B, B // 2, int(rng.integers(0, 3 * B + 1))]))
            extra = int(rng.integers(0, 5))  # input may be longer than output
            x = oracle.gen_noise(seed, pos, n + extra)
            pos += n
            yg, yo = np.zeros(n, np.float32), np.zeros(n, np.float32)
            g.process(x, yg)
            o.process(x, yo)
            run.add(yg, yo)
        elif op == "update":
            ln = int(rng.choice([L, int(rng.integers(0, L + 1))]))
            h = oracle.gen_ir(seed, upd, ln) if ln else np.zeros(0, np.float32)
            upd += 1
            g.update(h)
            o.update(h)
        elif op == "reset":
            g.reset()
            o.reset()
        else:
            g, o = g.clone(), o.clone()
        assert (g.current, g.fill, g.active_seg_count) == (o.current, o.fill, o.active_seg_count), (seed, step, op)
    run.check(TOL, f"FFTConvolver random ops, seed {seed}")


@pytest.mark.parametrize("seed", range(6))
def test_twostage_random_ops(F, seed):
    rng = np.random.default_rng(2000 + seed)
    H = int(2 ** rng.integers(3, 8))
    L = int(rng.integers(H, 60 * H))
    h = oracle.gen_ir(seed, 0, L)
    g = F.TwoStageFFTConvolver.init(h, H, L, async_tail=bool(seed % 2))
    o = oracle.TwoStageFFTConvolver.init(h, H, L)
    assert g.tail_block_size == o.tail_block_size
    pos, run = 0, WholeRun()
    for step in range(150):
        op = rng.choice(["process"] * 20 + ["reset", "clone"])
        if op == "process":
            n = int(rng.choice([H, H, H, int(rng.integers(0, H + 1))]))
            x = oracle.gen_noise(seed, pos, n)
            pos += n
            yg, yo = np.zeros(n, np.float32), np.zeros(n, np.float32)
            g.process(x, yg)
            o.process(x, yo)
            run.add(yg, yo)
        elif op == "reset":
            g.reset()
            o.reset()
        else:
            g, o = g.clone(), o.clone()
    run.check(TOL, f"TwoStage random ops, seed {seed}")


@pytest.mark.parametrize("seed", range(6))
def test_crossfade_random_ops(F, seed):
    rng = np.random.default_rng(3000 + seed)
    B = int(2 ** rng.integers(4, 9))
    L = int(rng.integers(B, 6 * B))
    fade = int(rng.integers(1, 5 * B))
    h = oracle.gen_ir(seed, 0, L)
    g = F.CrossfadeConvolver.new(F.FFTConvolver.init(h, B, L), L, B, fade)
    o = oracle.CrossfadeConvolver.new(oracle.FFTConvolver.init(h, B, L), L, B, fade)
    pos, upd, run = 0, 1, WholeRun()
    for step in range(90):
        if rng.random() < 0.15:
            ln = int(rng.choice([L, int(rng.integers(1, L + 1))]))
            hn = oracle.gen_ir(seed, upd, ln)
            upd += 1
            g.update(hn)
            o.update(hn)
        n_out = int(rng.choice([B, B, B, int(rng.integers(0, B + 1))]))
        x = oracle.gen_noise(seed, pos, B)  # both convolvers always consume max_buffer_size samples
        pos += B
        yg, yo = np.zeros(n_out, np.float32), np.zeros(n_out, np.float32)
        g.process(x, yg)
        o.process(x, yo)
        run.add(yg, yo)
        cnt, mix, appr, tgt = g.state()
        s = o.crossfader
        assert (cnt, appr, tgt, np.float32(mix)) == (s.counter, bool(s.approaching), s.target, np.float32(s.mix_value))
    run.check(TOL, f"Crossfade random ops, seed {seed}")
