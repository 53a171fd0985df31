"""K4 (mimo_tc.cuh) walks the ring in 16-slot blocks and pairs every block with an IR position; slots that do not
belong to an owned segment must meet the 16 zero positions in front of an IR row or fall off its end (TMA zero
fill), and every TMA box must start on an even IR position.  This checks the real stage enumeration (host copy of
the device code, no GPU needed) against the definition: segment i of [seg_lo, seg_hi) sits in slot (current + i) % S."""
import ctypes as C
import itertools

import numpy as np

from fft_convolution_b200 import _lib

KSEG, LEAD = 16, 16


def stages(S, cur, lo, hi):
    lib = _lib.load()
    n_max = 2 * (S // KSEG + 3)
    blk, cp, pos = (np.zeros(n_max, np.int32) for _ in range(3))
    n = lib.fcb_debug_tc_stages(S, cur, lo, hi, n_max, blk.ctypes.data_as(C.c_void_p), cp.ctypes.data_as(C.c_void_p),
                                pos.ctypes.data_as(C.c_void_p))
    assert 0 <= n <= n_max
    return [(int(blk[i]), int(cp[i]), int(pos[i])) for i in range(n)]


def check(S, cur, lo, hi):
    rows = hi - lo
    hits = {}
    for blk, copy, pos0 in stages(S, cur, lo, hi):
        assert pos0 % 2 == 0 and pos0 >= 0, (S, cur, lo, hi, pos0)        # 16-byte aligned box start
        assert blk >= 0 and copy in (0, 1)
        for j in range(KSEG):
            slot, pos = blk * KSEG + j, pos0 + j
            r = pos - LEAD - copy                                          # IR row at this position of this copy
            ring_zero = slot >= S                                          # ring slots past S are never written
            ir_zero = r < 0 or r >= rows                                   # lead pad / out of bounds
            if ring_zero or ir_zero:
                continue
            seg = lo + r
            assert (cur + seg) % S == slot, (S, cur, lo, hi, blk, copy, pos0, j)   # the right spectrum meets the right IR row
            hits[seg] = hits.get(seg, 0) + 1
    assert hits == {i: 1 for i in range(lo, hi)}, (S, cur, lo, hi)         # every owned segment exactly once


def test_every_ring_position_small_rings():
    for S in list(range(1, 40)) + [47, 48, 49, 63, 64, 65]:
        for cur in range(S):
            check(S, cur, 0, S)


def test_shards_and_odd_sizes():
    rng = np.random.default_rng(5)
    for _ in range(3000):
        S = int(rng.integers(1, 1000))
        cur = int(rng.integers(0, S))
        lo = int(rng.integers(0, S))
        hi = int(rng.integers(lo, S + 1))
        check(S, cur, lo, hi)


def test_headline_shape_all_positions():
    for cur in itertools.chain(range(0, 938, 7), (936, 937)):
        check(938, cur, 0, 938)
        check(938, cur, 469, 938)
