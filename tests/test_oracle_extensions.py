"""The oracle's EXTENSIONS beyond the reference (the two methods the reference leaves todo!() and the nested partition)
are pinned on the CPU by properties that follow from the reference's own semantics — they are what the CUDA path is
held to in tests/test_gpu_realtime.py."""
import numpy as np
import pytest

import oracle
from oracle import oracle_np
from refsignals import rms


def _run(conv, x, n):
    y = np.zeros_like(x)
    blk = np.zeros(n, np.float32)
    for i in range(x.size // n):
        conv.process(x[i * n:(i + 1) * n], blk)
        y[i * n:(i + 1) * n] = blk
    return y


@pytest.mark.parametrize("head,L", [(64, 12000), (32, 700), (128, 300)])
def test_twostage_update_on_a_fresh_convolver_equals_init(head, L):
    """nothing has been heard yet, so update(h1) must leave exactly the convolver init(h1) builds"""
    h0, h1 = oracle.gen_ir(1, 0, L), oracle.gen_ir(1, 1, L - 17)
    x = oracle.gen_noise(1, 0, head * 120)
    a = oracle.TwoStageFFTConvolver.init(h0, head, L)
    a.update_ext(h1)
    b = oracle.TwoStageFFTConvolver.init(h1, head, L)
    assert np.array_equal(_run(a, x, head), _run(b, x, head))


def test_twostage_update_keeps_the_strict_answer_and_panics_like_fftconvolver():
    a = oracle.TwoStageFFTConvolver.init(oracle.gen_ir(1, 0, 500), 32, 500)
    with pytest.raises(oracle.OraclePanic):
        a.update(np.zeros(10, np.float32))      # the reference: todo!()
    with pytest.raises(oracle.OraclePanic):
        a.update_ext(np.zeros(501, np.float32))  # longer than max_response_length (src/fft_convolver.rs:177-179)


def test_twostage_update_converges_to_the_new_response():
    """after max_response_length samples of new input nothing computed with the old response is left"""
    head, L = 64, 3000
    h0, h1 = oracle.gen_ir(2, 0, L), oracle.gen_ir(2, 1, L)
    x = oracle.gen_noise(2, 0, head * 200)
    a = oracle.TwoStageFFTConvolver.init(h0, head, L)
    ya = np.zeros_like(x)
    blk = np.zeros(head, np.float32)
    for i in range(200):
        if i == 40:
            a.update_ext(h1)
        a.process(x[i * head:(i + 1) * head], blk)
        ya[i * head:(i + 1) * head] = blk
    # a convolver that had h1 all along, fed the same input: equal once 40 blocks + L samples (+ two tail periods) have passed
    yb = _run(oracle.TwoStageFFTConvolver.init(h1, head, L), x, head)
    start = 40 * head + L + 3 * a.tail_block_size
    assert start < x.size - 10 * head
    assert np.max(np.abs(ya[start:] - yb[start:])) <= 1e-5 * rms(yb)


def test_crossfade_reset_lands_on_the_response_asked_for_last():
    B, L, fade = 64, 400, 300
    h0, h1 = oracle.gen_ir(3, 0, L), oracle.gen_ir(3, 1, L)
    x = oracle.gen_noise(3, 0, B * 40)
    c = oracle.CrossfadeConvolver.new(oracle.FFTConvolver.init(h0, B, L), L, B, fade)
    blk = np.zeros(B, np.float32)
    for i in range(6):
        if i == 3:
            c.update(h1)  # fade towards h1 starts
        c.process(x[i * B:(i + 1) * B], blk)
    assert c.is_crossfading()
    with pytest.raises(oracle.OraclePanic):
        c.reset()          # the reference: todo!()
    c.reset_ext()
    assert not c.is_crossfading()
    tail = x[6 * B:]
    fresh = oracle.FFTConvolver.init(h1, B, L)  # all audio forgotten, h1 is what is heard
    assert np.array_equal(_run(c, tail, B), _run(fresh, tail, B))


@pytest.mark.parametrize("head,L,stages", [(32, 40000, 3), (32, 40000, 4), (16, 9000, 5)])
def test_nested_partition_equals_the_uniform_convolver(head, L, stages):
    """the reference's twostage_equal test (src/tests.rs:148-175), one level deeper"""
    h = oracle.gen_ir(4, 0, L)
    n = 3 * 8192 // head
    x = oracle.gen_noise(4, 0, head * n)
    nested = oracle.TwoStageFFTConvolver.init(h, head, L, stages=stages, max_block=16384)
    blocks = nested.stage_blocks
    assert len(blocks) == stages + 0 or len(blocks) >= 3
    assert all(b2 >= b1 for b1, b2 in zip(blocks, blocks[1:]))
    y = _run(nested, x, head)
    yu = _run(oracle.FFTConvolver.init(h, head, L), x, head)
    yt = oracle_np.truth_f64(x, h)
    assert np.max(np.abs(y - yu)) <= 1e-5 * rms(yt) and np.max(np.abs(y - yt)) <= 1e-5 * rms(yt)
    # reset, rerun: the same output (the reference's reset_twostagefftconvolver, src/tests.rs:218-257)
    nested.reset()
    assert np.array_equal(_run(nested, x, head), y)
