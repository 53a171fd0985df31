"""bench.py's numpy generators must be the same data the oracle's C generators produce."""
import numpy as np

import bench
import oracle


def test_noise_matches_oracle():
    x = bench.synth_noise(5, 3, 100, 777)
    for c in range(3):
        assert np.array_equal(x[c], oracle.gen_noise(5 + c, 100, 777))


def test_irs_match_oracle():
    h = bench.synth_irs(2, 3, 4, 4800, chunk=2)
    for c in range(3):
        ref = oracle.gen_ir(2 + c, 4, 4800)
        assert np.max(np.abs(h[c] - ref)) <= 1e-9  # f64 sum order may differ in the last ulp of the norm
