"""Fixtures shared by the restated reference tests.

generate_sinusoid follows /root/reference/src/tests.rs:9-16: everything in f32,
`gain * (2.0 * PI * frequency * i as f32 / sample_rate).sin()`, evaluated left to right.
"""
import numpy as np

SAMPLE_RATE = np.float32(44100.0)


def generate_sinusoid(length: int, frequency: float, sample_rate=SAMPLE_RATE, gain: float = 1.0) -> np.ndarray:
    f = np.float32
    i = np.arange(length, dtype=np.float32)
    arg = ((f(2.0) * f(np.pi)) * f(frequency)).astype(np.float32) * i
    arg = (arg / f(sample_rate)).astype(np.float32)
    return (f(gain) * np.sin(arg, dtype=np.float32)).astype(np.float32)


def rms(x) -> float:
    x = np.asarray(x, dtype=np.float64)
    return float(np.sqrt(np.mean(x * x)))
