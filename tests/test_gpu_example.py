"""examples/compare_partitioned.py = the reference's demo (examples/compare_partitioned.rs)."""
import subprocess
import sys
import wave
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_compare_partitioned_example(tmp_path):
    r = subprocess.run([sys.executable, str(ROOT / "examples" / "compare_partitioned.py"), "--outdir", str(tmp_path)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Uniform took" in r.stdout and "Partitioned took" in r.stdout
    rel = float(r.stdout.split("max abs diff vs block by block: ")[1].split()[0])
    assert rel <= 1e-5  # one multi-block call vs block by block (split delay lines): same arithmetic, other association
    diff = float(r.stdout.split("max_abs_diff = ")[1].split()[0])
    assert diff < 1e-4  # 128 000-tap sinusoid IR: two f32 partitionings of a large-gain filter
    for name in ("output_a.wav", "output_b.wav"):
        with wave.open(str(tmp_path / name), "rb") as w:
            assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 44100, 64000)
            a = np.frombuffer(w.readframes(64000), dtype="<i2")
            assert np.abs(a).max() > 0
