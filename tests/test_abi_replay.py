"""Layer 1 of the C ABI driven by a plain C host (tests/abi/replay.c) that issues the exact call sequence of
the Rust binding's `CudaFFTConvolver` — push_input -> fft_forward -> mac -> ifft_ola -> fetch with `current` and
`input_buffer_fill` held by the caller, plus the multi-block branch.  No Rust toolchain exists in this image, so
this is the test of that sequence that does not go through the C++ host mirror."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle
import refgolden
from refsignals import rms

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "fft_convolution_b200"


@pytest.fixture(scope="module")
def replay_bin(tmp_path_factory):
    from fft_convolution_b200 import build
    build.build()
    exe = tmp_path_factory.mktemp("abi") / "replay"
    subprocess.run(["/usr/bin/gcc", "-O2", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "abi" / "replay.c"),
                    "-o", str(exe), f"-L{PKG}", "-lfftconv_b200", f"-Wl,-rpath,{PKG}", "-lm"], check=True)
    return exe


def test_replay_harness_compiles_and_links_against_the_header(replay_bin):
    """every layer-1 symbol the Rust binding declares resolves from the header + the shared library (no GPU needed)"""
    assert replay_bin.exists()
    out = subprocess.run(["nm", "-u", str(replay_bin)], capture_output=True, text=True, check=True).stdout
    for sym in ("fcb_engine_create", "fcb_engine_set_ir", "fcb_engine_push_input", "fcb_engine_fft_forward", "fcb_engine_mac",
                "fcb_engine_ifft_ola", "fcb_engine_fetch", "fcb_engine_scratch", "fcb_engine_process_blocks", "fcb_engine_reset"):
        assert sym in out, sym


def _run(replay_bin, tmp_path, case):
    assert case["kind"] == "uniform"
    for i, h in enumerate(case["irs"]):
        np.asarray(h, "<f4").tofile(tmp_path / f"h{i}.f32")
    np.asarray(case["x"], "<f4").tofile(tmp_path / "x.f32")
    upd = case["updates"][0][0] if case["updates"] else -1
    rst = case["resets"][0] if case["resets"] else -1
    r = subprocess.run([str(replay_bin), str(tmp_path), str(case["block"]), str(case["max_len"]), str(upd), str(rst),
                        *[str(s) for s in case["sizes"]]], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return np.fromfile(tmp_path / "y.f32", "<f4")


UNIFORM = [c for c in refgolden.dump_cases() if c["kind"] == "uniform" and c["max_len"] <= 12000]


@pytest.mark.gpu
@pytest.mark.parametrize("case", UNIFORM, ids=lambda c: c["name"])
def test_layer1_sequence_matches_oracle_and_goldens(replay_bin, tmp_path, case):
    y = _run(replay_bin, tmp_path, case)
    yo = refgolden.replay(case, oracle)
    assert np.max(np.abs(y - yo)) <= 1e-5 * max(rms(yo), 1e-3), case["name"]
    g = refgolden.GOLDEN_DIR / f"{case['name']}.npz"
    if g.exists():
        assert np.max(np.abs(y - np.load(g)["y"])) <= 1e-5 * rms(yo)
    ref = refgolden.GOLDEN_DIR / f"ref_{case['name']}.bin"  # reference-made vectors, once they exist
    if ref.exists():
        assert np.max(np.abs(y - refgolden.read_case(ref)["y"])) <= 1e-5 * max(rms(yo), 1e-3)


@pytest.mark.gpu
@pytest.mark.usefixtures("exact_paths")
def test_layer1_sequence_equals_the_host_mirror_bit_for_bit(replay_bin, tmp_path):
    """ragged calls (partial blocks: K1 + K3 per chunk, K2 only at block start) and one multi-block call"""
    import fft_convolution_b200 as F
    case = dict(refgolden.dump_cases()[1])  # uniform_b256_l3000_ragged
    case["sizes"] = [100, 256, 37, 300, 1024]
    assert np.array_equal(_run(replay_bin, tmp_path, case), refgolden.replay(case, F))
