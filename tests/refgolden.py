"""Reference-made golden vectors: reader / writer of the `ref_<case>.bin` container that
rust/golden_dump writes, the Python twin of its case list, and the generic replay.

A case = one convolver (`kind` + constructor arguments), a cyclic list of call sizes, `update`
calls placed before given process() calls, `reset` calls likewise (rust/golden_dump/src/main.rs).
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

import oracle
from refsignals import generate_sinusoid

MAGIC = b"FCBREF1\0"
GOLDEN_DIR = Path(__file__).resolve().parent / "golden"
DUMP_MAIN = Path(__file__).resolve().parent.parent / "rust" / "golden_dump" / "src" / "main.rs"


def read_case(path) -> dict:
    raw = Path(path).read_bytes()
    assert raw[:8] == MAGIC, f"{path}: not a golden_dump file"
    (n,), p = struct.unpack_from("<I", raw, 8), 12
    out = {}
    for _ in range(n):
        (ln,) = struct.unpack_from("<I", raw, p); p += 4
        name = raw[p:p + ln].decode(); p += ln
        dtype, count = struct.unpack_from("<IQ", raw, p); p += 12
        if dtype == 1:
            out[name] = raw[p:p + count].decode(); p += count
        else:
            out[name] = np.frombuffer(raw, dtype="<f4", count=count, offset=p).copy(); p += 4 * count
    meta = dict(kv.split("=", 1) for kv in out.pop("meta").split(";"))
    ints = lambda s: [int(v) for v in s.split(",") if v]  # noqa: E731
    case = dict(name=Path(path).stem[len("ref_"):], kind=meta["kind"], block=int(meta["block"]), max_len=int(meta["max_len"]),
                xf_len=int(meta["xf_len"]), xf_buf=int(meta["xf_buf"]), fade=int(meta["fade"]), sizes=ints(meta["sizes"]),
                updates=[tuple(int(a) for a in u.split(":")) for u in meta["updates"].split(",") if u],
                resets=ints(meta["resets"]), irs=[out[f"h{i}"] for i in range(int(meta["n_irs"]))], x=out["x"])
    case["y"] = out["y"]
    return case


def write_case(directory, case: dict, y: np.ndarray) -> Path:
    """The Rust writer's format, byte for byte (used to test the reader and the replay without cargo)."""
    meta = "kind={kind};block={block};max_len={max_len};xf_len={xf_len};xf_buf={xf_buf};fade={fade}".format(**case)
    meta += ";sizes=" + ",".join(str(v) for v in case["sizes"])
    meta += ";updates=" + ",".join(f"{c}:{i}" for c, i in case["updates"])
    meta += ";resets=" + ",".join(str(v) for v in case["resets"]) + f";n_irs={len(case['irs'])}"
    blob = bytearray(MAGIC) + struct.pack("<I", 3 + len(case["irs"]))

    def entry(name, dtype, payload, count):
        nonlocal blob
        blob += struct.pack("<I", len(name)) + name.encode() + struct.pack("<IQ", dtype, count) + payload

    entry("meta", 1, meta.encode(), len(meta.encode()))
    for i, h in enumerate(case["irs"]):
        entry(f"h{i}", 0, np.asarray(h, "<f4").tobytes(), len(h))
    entry("x", 0, np.asarray(case["x"], "<f4").tobytes(), len(case["x"]))
    entry("y", 0, np.asarray(y, "<f4").tobytes(), len(y))
    path = Path(directory) / f"ref_{case['name']}.bin"
    path.write_bytes(bytes(blob))
    return path


def _case(name, kind, block, max_len, sizes, irs, x, *, updates=(), resets=(), xf=(0, 0, 0)) -> dict:
    return dict(name=name, kind=kind, block=block, max_len=max_len, xf_len=xf[0], xf_buf=xf[1], fade=xf[2], sizes=list(sizes),
                updates=list(updates), resets=list(resets), irs=[np.asarray(h, np.float32) for h in irs], x=np.asarray(x, np.float32))


def dump_cases() -> list[dict]:
    """Python twin of rust/golden_dump/src/main.rs::cases() — same names, seeds, shapes and schedules."""
    g, n, s = oracle.gen_ir, oracle.gen_noise, generate_sinusoid
    v = [
        _case("uniform_b64_l1000", "uniform", 64, 1000, [64], [g(11, 0, 1000)], n(11, 0, 64 * 24)),
        _case("uniform_b256_l3000_ragged", "uniform", 256, 3000, [100, 256, 37, 300], [g(11, 0, 3000)], n(11, 0, 256 * 10)),
        _case("uniform_b512_l5000", "uniform", 512, 5000, [512], [g(11, 0, 5000)], n(11, 0, 512 * 8)),
        _case("twostage_h64_l12000", "twostage", 64, 12000, [64], [g(12, 0, 12000)], n(12, 0, 64 * 80)),
        _case("crossfade_b64_l300_update6", "crossfade_new", 64, 300, [64], [g(13, 0, 300), g(13, 1, 300)], n(13, 0, 64 * 24),
              updates=[(6, 1)], xf=(300, 64, 200)),
        _case("cfg0_uniform_b256_l48000", "uniform", 256, 48000, [256], [g(0, 0, 48000)], n(0, 0, 256 * 400)),
        _case("cfg1_twostage_h128_l240000", "twostage", 128, 240000, [128], [g(0, 0, 240000)], n(0, 0, 128 * 200)),
        _case("cfg2_crossfade_init_b512_l96000", "crossfade_init", 512, 96000, [512], [g(0, u, 96000) for u in range(3)],
              n(0, 0, 512 * 260), updates=[(50, 1), (100, 2), (150, 1), (200, 2)]),
        _case("cfg3_uniform_b512_l96000", "uniform", 512, 96000, [512], [g(7, 0, 96000)], n(7, 0, 512 * 220)),
    ]
    a, b, x = s(512, 1000.0, gain=1.0), s(512, 2000.0, gain=0.7), s(16 * 512, 1300.0, gain=1.0)
    v.append(_case("reftest_update_is_reset", "uniform", 512, 512, [512], [a, b], x, updates=[(8, 1)]))
    v.append(_case("reftest_crossfade_convolver", "crossfade_new", 512, 512, [512], [a, b], x, updates=[(8, 1)], xf=(512, 512, 512)))
    h, x = s(128, 1000.0, gain=0.1), s(200 * 128, 1300.0, gain=1.0)
    v.append(_case("reftest_block_size_equal_b64", "uniform", 64, 128, [128], [h], x))
    v.append(_case("reftest_block_size_equal_b128", "uniform", 128, 128, [128], [h], x))
    h, x = s(12000, 1000.0, gain=0.1), s(300 * 64, 1300.0, gain=1.0)
    v.append(_case("reftest_twostage_equal_uniform_b32", "uniform", 32, 12000, [64], [h], x))
    v.append(_case("reftest_twostage_equal_twostage_h64", "twostage", 64, 12000, [64], [h], x))
    x1 = s(300 * 64, 1300.0, gain=0.1)
    x = np.concatenate([x1, x1])
    v.append(_case("reftest_reset_uniform_b64", "uniform", 64, 12000, [64], [h], x, resets=[300]))
    v.append(_case("reftest_reset_twostage_h64", "twostage", 64, 12000, [64], [h], x, resets=[300]))
    d = np.zeros(1024, np.float32)
    d[0] = 1.0
    ones = np.ones(1024, np.float32)
    v.append(_case("reftest_passthrough_uniform", "uniform", 1024, 1024, [1024], [d], ones))
    v.append(_case("reftest_passthrough_twostage", "twostage", 1024, 1024, [1024], [d], ones))
    v.append(_case("reftest_passthrough_crossfade", "crossfade_new", 1024, 1024, [1024], [d], ones, xf=(1024, 1024, 1024)))
    return v


def replay(case: dict, impl) -> np.ndarray:
    """Run `case` through `impl` (the `oracle` package or `fft_convolution_b200`): same calls, same order."""
    h0, B, L = case["irs"][0], case["block"], case["max_len"]
    if case["kind"] == "uniform":
        conv = impl.FFTConvolver.init(h0, B, L)
    elif case["kind"] == "twostage":
        conv = impl.TwoStageFFTConvolver.init(h0, B, L)
    elif case["kind"] == "crossfade_new":
        conv = impl.CrossfadeConvolver.new(impl.FFTConvolver.init(h0, B, L), case["xf_len"], case["xf_buf"], case["fade"])
    else:
        conv = impl.CrossfadeConvolver.init(h0, B, L)
    x = case["x"]
    y = np.zeros_like(x)
    p = call = 0
    while p < x.size:
        for at, which in case["updates"]:
            if at == call:
                conv.update(case["irs"][which])
        if call in case["resets"]:
            conv.reset()
        n = min(case["sizes"][call % len(case["sizes"])], x.size - p)
        out = np.zeros(n, np.float32)
        conv.process(np.ascontiguousarray(x[p:p + n]), out)
        y[p:p + n] = out
        p += n
        call += 1
    return y
