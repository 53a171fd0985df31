"""Host logic of the CUDA-core matrix kernel's launch plan: the segment range is cut into z chunks so that the
grid fills whole waves of 148 CTAs (one CTA is resident per SM).  No GPU needed."""
import ctypes as C
import math

import numpy as np

from fft_convolution_b200 import _lib


def plan(logb, n_in, n_out, ns, nsegs):
    z, zl = C.c_int(), C.c_int()
    rc = _lib.load().fcb_debug_mac_tile_plan(logb, n_in, n_out, ns, nsegs, C.byref(z), C.byref(zl))
    return rc, z.value, zl.value


def base_ctas(logb, n_in, n_out, ns):
    B = 1 << logb
    tiles = 1 if B < 512 else B // 512
    wide = ns >= 2
    ot, st = (4, 4) if wide else (8, 1)
    return tiles * math.ceil(n_out / ot) * n_in * math.ceil(ns / st)


def test_chunks_cover_the_segment_range():
    rng = np.random.default_rng(0)
    for _ in range(2000):
        logb = int(rng.integers(6, 15))
        n_in, n_out, ns = (int(rng.integers(1, 33)) for _ in range(3))
        nsegs = int(rng.integers(0, 3000))
        rc, z, zl = plan(logb, n_in, n_out, ns, nsegs)
        assert rc == 0 and z >= 1 and zl >= 1
        assert z * zl >= nsegs                      # every segment belongs to a chunk
        assert nsegs == 0 or (z - 1) * zl < nsegs   # and no chunk is empty


def test_headline_matrix_fills_whole_waves():
    # 16 x 16, one stream, block 512, 937 segments: 32 CTAs per chunk; 19 chunks were 4.1 waves
    rc, z, zl = plan(9, 16, 16, 1, 937)
    assert rc == 0
    waves = 32 * z / 148
    assert waves / math.ceil(waves) >= 0.95
    assert 12 <= z <= 26


def test_small_blocks_are_left_to_the_generic_kernel():
    assert plan(5, 4, 4, 1, 100)[0] != 0
