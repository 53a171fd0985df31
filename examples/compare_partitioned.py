#!/usr/bin/env python
"""The reference's only demo/benchmark (examples/compare_partitioned.rs:9-68) on the B200 engine:
1000 blocks of 64 samples through FFTConvolver (uniform) and TwoStageFFTConvolver with a
128 000-tap sinusoid IR at 44.1 kHz; prints both wall times and the max-abs-diff, and writes the
two outputs as 16-bit mono WAV files (examples/util/mod.rs:21-40).  Needs a CUDA device."""
import argparse
import sys
import time
import wave
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fft_convolution_b200 as F  # noqa: E402

SAMPLE_RATE = 44100


def generate_sinusoid(num_samples: int, freq: float, sample_rate: int, gain: float) -> np.ndarray:
    """examples/util/mod.rs:7-19: evaluated in f64, stored as f32"""
    t = np.arange(num_samples, dtype=np.float64) / float(sample_rate)
    return (gain * np.sin(2.0 * np.pi * freq * t)).astype(np.float32)


def save_wav(filename: str, samples: np.ndarray, sample_rate: int) -> None:
    """examples/util/mod.rs:21-40: `(sample * i16::MAX as f32) as i16` (truncating, saturating)"""
    scaled = np.clip(np.trunc(samples.astype(np.float32) * np.float32(32767.0)), -32768, 32767).astype("<i2")
    with wave.open(filename, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sample_rate)
        w.writeframes(scaled.tobytes())
    print(f"Saved: {filename}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--outdir", default=".")
    a = ap.parse_args()
    block_size, n_blocks, response_len = 64, 1000, 128_000
    response = generate_sinusoid(response_len, 1000.0, SAMPLE_RATE, 0.1)
    convolver_a = F.FFTConvolver.init(response, block_size, response.size)
    convolver_b = F.TwoStageFFTConvolver.init(response, block_size, response.size)
    output_a = np.zeros(block_size * n_blocks, np.float32)
    output_b = np.zeros(block_size * n_blocks, np.float32)
    block = np.zeros(block_size, np.float32)
    x = generate_sinusoid(n_blocks * block_size, 1300.0, SAMPLE_RATE, 0.1)

    t0 = time.perf_counter()
    for i in range(n_blocks):
        convolver_a.process(x[i * block_size:(i + 1) * block_size], block)
        output_a[i * block_size:(i + 1) * block_size] = block
    print(f"Uniform took = {(time.perf_counter() - t0) * 1000.0:.2f} ms")

    t0 = time.perf_counter()
    for i in range(n_blocks):
        convolver_b.process(x[i * block_size:(i + 1) * block_size], block)
        output_b[i * block_size:(i + 1) * block_size] = block
    print(f"Partitioned took = {(time.perf_counter() - t0) * 1000.0:.2f} ms (tail block {convolver_b.tail_block_size})")

    print(f"max_abs_diff = {float(np.max(np.abs(output_a - output_b)))!r}")

    # the same uniform convolver handed the whole input in ONE call: the engine batches the 1000 blocks over time
    # (csrc/offline_kernels.cuh).  Same arithmetic per block; the block-by-block run above cut each delay line over
    # several CTAs (one channel cannot fill the GPU otherwise), so the f32 sums are associated differently
    convolver_c = F.FFTConvolver.init(response, block_size, response.size)
    convolver_c.reserve(x.size)  # process() never allocates: size the multi-block workspace ahead of the call
    output_c = np.zeros_like(output_a)
    convolver_c.process(x, output_c)
    convolver_c.reset()
    t0 = time.perf_counter()
    convolver_c.process(x, output_c)
    rel = float(np.max(np.abs(output_a - output_c))) / float(np.sqrt(np.mean(output_a.astype(np.float64) ** 2)))
    print(f"Uniform, one call = {(time.perf_counter() - t0) * 1000.0:.2f} ms (max abs diff vs block by block: {rel:.2e} x RMS)")
    out = Path(a.outdir)
    save_wav(str(out / "output_a.wav"), output_a, SAMPLE_RATE)
    save_wav(str(out / "output_b.wav"), output_b, SAMPLE_RATE)


if __name__ == "__main__":
    main()
