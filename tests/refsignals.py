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


class WholeRun:
    """north_star's criterion over one whole run: max |got - want| <= tol x RMS of the entire reference output
    (per-block RMS over 32-64 samples is noise; the run's RMS is what "relative to the output RMS" means)."""

    def __init__(self):
        self.err, self.sq, self.n = 0.0, 0.0, 0

    def add(self, got, want):
        got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
        assert got.shape == want.shape
        if want.size:
            self.err = max(self.err, float(np.max(np.abs(got - want))))
            self.sq += float(np.sum(want * want))
            self.n += want.size

    def ratio(self) -> float:
        if self.n == 0 or self.sq == 0.0:
            return 0.0 if self.err == 0.0 else float("inf")
        return self.err / float(np.sqrt(self.sq / self.n))

    def check(self, tol: float = 1e-5, what=""):
        assert self.ratio() <= tol, f"{what}: max-abs error {self.ratio():.3e} x whole-run RMS exceeds {tol:g}"
