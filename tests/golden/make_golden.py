#!/usr/bin/env python
"""Generates tests/golden/*.npz: small seeded input/output vectors of the hot path.

The reference is Rust and cannot run in this environment (no cargo/rustc; its FFT is in the
absent realfft/rustfft crates), so these vectors come from the CPU oracle
(oracle/fftconv_oracle.c), accepted only where the independent numpy restatement
(oracle/oracle_np.py, pocketfft) and an f64 direct convolution agree to <= 1e-5 * RMS.
They pin the oracle and the CUDA path against silent drift; they are not reference bit patterns.

    python tests/golden/make_golden.py        (from the repo root)
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
from oracle import oracle_np  # noqa: E402

OUT = Path(__file__).resolve().parent


def run(conv, x, sizes):
    y = np.zeros_like(x)
    p = k = 0
    while p < x.size:
        n = min(sizes[k % len(sizes)], x.size - p)
        blk = np.zeros(n, np.float32)
        conv.process(x[p:p + n], blk)
        y[p:p + n] = blk
        p += n
        k += 1
    return y


def rms(v):
    return float(np.sqrt(np.mean(np.asarray(v, np.float64) ** 2)))


def main():
    cases = {}
    # uniform convolver: (block, ir_len, call sizes)
    for name, B, L, sizes, nblk in [("uniform_b64_l1000", 64, 1000, [64], 24), ("uniform_b256_l3000_ragged", 256, 3000, [100, 256, 37, 300], 10),
                                    ("uniform_b512_l5000", 512, 5000, [512], 8)]:
        h, x = oracle.gen_ir(11, 0, L), oracle.gen_noise(11, 0, B * nblk)
        y = run(oracle.FFTConvolver.init(h, B, L), x, sizes)
        yn = run(oracle_np.FFTConvolverNP.init(h, B, L), x, sizes)
        yt = oracle_np.truth_f64(x, h)
        assert max(np.max(np.abs(y - yt)), np.max(np.abs(yn - yt)), np.max(np.abs(y - yn))) <= 1e-5 * rms(yt), name
        cases[name] = dict(kind="uniform", block=B, ir_len=L, sizes=np.array(sizes), h=h, x=x, y=y)
    # two-stage: head 64, 12 000 taps => T = 1024 (16 + 16 + 10 segments)
    H, L = 64, 12000
    h, x = oracle.gen_ir(12, 0, L), oracle.gen_noise(12, 0, H * 80)
    y = run(oracle.TwoStageFFTConvolver.init(h, H, L), x, [H])
    yn = run(oracle_np.TwoStageNP.init(h, H, L), x, [H])
    yt = oracle_np.truth_f64(x, h)
    assert max(np.max(np.abs(y - yt)), np.max(np.abs(yn - yt))) <= 1e-5 * rms(yt)
    cases["twostage_h64_l12000"] = dict(kind="twostage", block=H, ir_len=L, sizes=np.array([H]), h=h, x=x, y=y)
    # crossfade: fade 200 + hold 64, update at block 6
    B, L = 64, 300
    h0, h1, x = oracle.gen_ir(13, 0, L), oracle.gen_ir(13, 1, L), oracle.gen_noise(13, 0, B * 24)
    xf = oracle.CrossfadeConvolver.new(oracle.FFTConvolver.init(h0, B, L), L, B, 200)
    y = np.zeros_like(x)
    blk = np.zeros(B, np.float32)
    for i in range(24):
        if i == 6:
            xf.update(h1)
        xf.process(x[i * B:(i + 1) * B], blk)
        y[i * B:(i + 1) * B] = blk
    cases["crossfade_b64_l300_update6"] = dict(kind="crossfade", block=B, ir_len=L, sizes=np.array([B]), h=h0, h1=h1, x=x, y=y,
                                               fade=np.array(200), update_block=np.array(6))
    for name, d in cases.items():
        np.savez_compressed(OUT / f"{name}.npz", **d)
        print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in d.items() if k != "kind"})


if __name__ == "__main__":
    main()
