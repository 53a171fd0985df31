"""ctypes binding of oracle/fftconv_oracle.c (TEST INFRASTRUCTURE ONLY).

The classes mirror the reference's `Convolution` trait (src/lib.rs:5-14): init / update /
reset / process, with the reference's panics surfaced as OraclePanic.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "liboracle.so"


class OraclePanic(RuntimeError):
    """The reference would panic!/assert!/todo!() at this point."""


def build(force: bool = False) -> Path:
    src = _HERE / "fftconv_oracle.c"
    hdr = _HERE / "fftconv_oracle.h"
    if force or not _SO.exists() or _SO.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    return _SO


_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_c64p = np.ctypeslib.ndpointer(dtype=np.complex64, flags="C_CONTIGUOUS")
_sz = C.c_size_t
_vp = C.c_void_p


class _CrossfaderStruct(C.Structure):
    _fields_ = [("fading_samples", C.c_int64), ("hold_samples", C.c_int64), ("counter", C.c_int64),
                ("mix_value_step", C.c_float), ("mix_value", C.c_float),
                ("approaching", C.c_int), ("target", C.c_int)]


class OracleLib:
    def __init__(self, path: Path):
        L = self.lib = C.CDLL(str(path))

        def sig(name, res, *args):
            f = getattr(L, name)
            f.restype = res
            f.argtypes = list(args)
            return f

        sig("orc_plan_new", _vp, _sz)
        sig("orc_plan_free", None, _vp)
        sig("orc_rfft_forward", None, _vp, _f32p, _c64p)
        sig("orc_rfft_inverse", None, _vp, _c64p, _f32p)
        sig("orc_complex_multiply_accumulate", None, _c64p, _c64p, _c64p, _sz)
        sig("orc_fftconv_init", _vp, _f32p, _sz, _sz, _sz)
        sig("orc_fftconv_default", _vp)
        sig("orc_fftconv_clone", _vp, _vp)
        sig("orc_fftconv_free", None, _vp)
        sig("orc_fftconv_update", C.c_int, _vp, _f32p, _sz)
        sig("orc_fftconv_reset", None, _vp)
        sig("orc_fftconv_process", C.c_int, _vp, _f32p, _sz, _f32p, _sz)
        for n in ("block_size", "seg_count", "active_seg_count", "current", "fill"):
            sig(f"orc_fftconv_{n}", _sz, _vp)
        sig("orc_fftconv_segment_ir", _vp, _vp, _sz)
        sig("orc_fftconv_segment", _vp, _vp, _sz)
        sig("orc_fftconv_premul", _vp, _vp)
        sig("orc_fftconv_overlap", _vp, _vp)
        sig("orc_compute_tail_block_size", _sz, _sz, _sz)
        sig("orc_twostage_init", _vp, _f32p, _sz, _sz, _sz)
        sig("orc_twostage_init_tail", _vp, _f32p, _sz, _sz, _sz, _sz)
        sig("orc_twostage_init_multi", _vp, _f32p, _sz, _sz, _sz, _sz, _sz)
        sig("orc_twostage_stage_blocks", _sz, _vp, C.POINTER(_sz), _sz)
        sig("orc_twostage_clone", _vp, _vp)
        sig("orc_twostage_free", None, _vp)
        sig("orc_twostage_update", C.c_int, _vp, _f32p, _sz)
        sig("orc_twostage_update_ext", C.c_int, _vp, _f32p, _sz)
        sig("orc_twostage_reset", None, _vp)
        sig("orc_twostage_process", C.c_int, _vp, _f32p, _sz, _f32p, _sz)
        sig("orc_twostage_tail_block_size", _sz, _vp)
        sig("orc_crossfader_new", None, C.POINTER(_CrossfaderStruct), _sz, _sz)
        sig("orc_crossfader_fade_into", None, C.POINTER(_CrossfaderStruct), C.c_int)
        sig("orc_crossfader_mix", C.c_float, C.POINTER(_CrossfaderStruct), C.c_float, C.c_float)
        sig("orc_raised_cosine_mix", C.c_float, C.c_float, C.c_float, C.c_float)
        sig("orc_crossfade_new", _vp, _vp, _sz, _sz, _sz)
        sig("orc_crossfade_init", _vp, _f32p, _sz, _sz, _sz)
        sig("orc_crossfade_free", None, _vp)
        sig("orc_crossfade_update", C.c_int, _vp, _f32p, _sz)
        sig("orc_crossfade_process", C.c_int, _vp, _f32p, _sz, _f32p, _sz)
        sig("orc_crossfade_reset", C.c_int, _vp)
        sig("orc_crossfade_reset_ext", C.c_int, _vp)
        sig("orc_crossfade_is_crossfading", C.c_int, _vp)
        sig("orc_crossfade_crossfader", C.POINTER(_CrossfaderStruct), _vp)
        sig("orc_mix64", C.c_uint64, C.c_uint64)
        sig("orc_gen_noise", None, _f32p, C.c_uint64, _sz, _sz)
        sig("orc_gen_ir", None, _f32p, C.c_uint64, C.c_uint64, _sz)
        sig("orc_direct_conv_f64", None, _f32p, _sz, _f32p, _sz, _f64p)
        sig("orc_batch_fftconv_run", C.c_double, _sz, _sz, _sz, _f32p, _f32p, _f32p, _sz, _sz, C.c_int)
        sig("orc_max_threads", C.c_int)
        sig("orc_batch_twostage_run", C.c_double, _sz, _sz, _sz, _sz, _f32p, _vp, _sz, _sz, _f32p, _f32p, _sz, _sz, C.c_int)
        sig("orc_batch_crossfade_run", C.c_double, _sz, _sz, _sz, _f32p, _vp, _sz, _sz, _f32p, _f32p, _sz, C.c_int)


_LIB: OracleLib | None = None


def load() -> OracleLib:
    global _LIB
    if _LIB is None:
        try:
            _LIB = OracleLib(build())
        except (OSError, subprocess.CalledProcessError):
            _LIB = OracleLib(build(force=True))
    return _LIB


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _check(rc: int, what: str):
    if rc != 0:
        raise OraclePanic(what)


class FFTConvolver:
    """Oracle mirror of the reference FFTConvolver (src/fft_convolver.rs:86-307)."""

    def __init__(self, handle, lib: OracleLib):
        self._h, self._lib = handle, lib

    @classmethod
    def init(cls, response, block_size: int, max_response_length: int) -> "FFTConvolver":
        lib = load()
        r = _f32(response)
        h = lib.lib.orc_fftconv_init(r, r.size, block_size, max_response_length)
        if not h:
            raise OraclePanic("max_response_length must be at least the length of the initial impulse response")
        return cls(h, lib)

    def clone(self) -> "FFTConvolver":
        return FFTConvolver(self._lib.lib.orc_fftconv_clone(self._h), self._lib)

    def update(self, response):
        r = _f32(response)
        _check(self._lib.lib.orc_fftconv_update(self._h, r, r.size),
               "New impulse response is longer than initialized length")

    def reset(self):
        self._lib.lib.orc_fftconv_reset(self._h)

    def process(self, input, output):
        x = _f32(input)
        assert output.dtype == np.float32 and output.flags.c_contiguous
        _check(self._lib.lib.orc_fftconv_process(self._h, x, x.size, output, output.size),
               "input shorter than output")

    # introspection -------------------------------------------------------------------------
    def _get(self, name):
        return getattr(self._lib.lib, f"orc_fftconv_{name}")(self._h)

    block_size = property(lambda s: s._get("block_size"))
    seg_count = property(lambda s: s._get("seg_count"))
    active_seg_count = property(lambda s: s._get("active_seg_count"))
    current = property(lambda s: s._get("current"))
    fill = property(lambda s: s._get("fill"))

    def _cpx(self, ptr, n):
        buf = (C.c_float * (2 * n)).from_address(ptr)
        return np.frombuffer(buf, dtype=np.complex64).copy()

    def segment_ir(self, i):
        return self._cpx(self._lib.lib.orc_fftconv_segment_ir(self._h, i), self.block_size + 1)

    def segment(self, i):
        return self._cpx(self._lib.lib.orc_fftconv_segment(self._h, i), self.block_size + 1)

    def premul(self):
        return self._cpx(self._lib.lib.orc_fftconv_premul(self._h), self.block_size + 1)

    def overlap(self):
        buf = (C.c_float * self.block_size).from_address(self._lib.lib.orc_fftconv_overlap(self._h))
        return np.frombuffer(buf, dtype=np.float32).copy()

    def _release(self):
        h, self._h = self._h, None
        return h

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.lib.orc_fftconv_free(self._h)
            self._h = None


class TwoStageFFTConvolver:
    """Oracle mirror of TwoStageFFTConvolver (src/fft_convolver.rs:323-526)."""

    def __init__(self, handle, lib):
        self._h, self._lib = handle, lib

    @classmethod
    def init(cls, response, block_size, max_response_length, forced_tail: int = 0, stages: int = 2, max_block: int = 0):
        """stages > 2: EXTENSION, the partition nested (the tail is again a two-stage convolver)"""
        lib = load()
        r = _f32(response)
        if stages > 2:
            h = lib.lib.orc_twostage_init_multi(r, r.size, block_size, max_response_length, stages, max_block)
        else:
            h = lib.lib.orc_twostage_init_tail(r, r.size, block_size, max_response_length, forced_tail)
        if not h:
            raise OraclePanic("max_response_length must be at least the length of the initial impulse response")
        return cls(h, lib)

    def clone(self):
        return TwoStageFFTConvolver(self._lib.lib.orc_twostage_clone(self._h), self._lib)

    def update(self, response):
        r = _f32(response)
        _check(self._lib.lib.orc_twostage_update(self._h, r, r.size), "not yet implemented")

    def update_ext(self, response):
        """EXTENSION beyond the reference (todo!() there): per-stage FFTConvolver::update on the re-sliced response."""
        r = _f32(response)
        _check(self._lib.lib.orc_twostage_update_ext(self._h, r, r.size),
               "New impulse response is longer than initialized length")

    def reset(self):
        self._lib.lib.orc_twostage_reset(self._h)

    def process(self, input, output):
        x = _f32(input)
        _check(self._lib.lib.orc_twostage_process(self._h, x, x.size, output, output.size),
               "assertion failed: input.len() <= self.head_block_size (or length mismatch)")

    @property
    def tail_block_size(self):
        return self._lib.lib.orc_twostage_tail_block_size(self._h)

    @property
    def stage_blocks(self):
        buf = (_sz * 16)()
        n = self._lib.lib.orc_twostage_stage_blocks(self._h, buf, 16)
        return [int(buf[i]) for i in range(min(n, 16))]

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.lib.orc_twostage_free(self._h)
            self._h = None


class Crossfader:
    """Oracle mirror of Crossfader<RaisedCosineMixer> (src/crossfade_convolver.rs:192-279)."""
    A, B = 0, 1

    def __init__(self, fading_samples, hold_samples):
        self._lib = load()
        self.s = _CrossfaderStruct()
        self._lib.lib.orc_crossfader_new(C.byref(self.s), fading_samples, hold_samples)

    def fade_into(self, target):
        self._lib.lib.orc_crossfader_fade_into(C.byref(self.s), target)

    def mix(self, a, b):
        return np.float32(self._lib.lib.orc_crossfader_mix(C.byref(self.s), a, b))

    @property
    def state(self):
        return ("Approaching" if self.s.approaching else "Reached", self.s.target)


class CrossfadeConvolver:
    """Oracle mirror of CrossfadeConvolver<FFTConvolver> (src/crossfade_convolver.rs:3-105)."""

    def __init__(self, handle, lib):
        self._h, self._lib = handle, lib

    @classmethod
    def new(cls, convolver: FFTConvolver, max_response_length, max_buffer_size, crossfade_samples):
        """Consumes `convolver` (the reference moves it in)."""
        lib = load()
        return cls(lib.lib.orc_crossfade_new(convolver._release(), max_response_length,
                                             max_buffer_size, crossfade_samples), lib)

    @classmethod
    def init(cls, response, max_block_size, max_response_length):
        lib = load()
        r = _f32(response)
        h = lib.lib.orc_crossfade_init(r, r.size, max_block_size, max_response_length)
        if not h:
            raise OraclePanic("max_response_length must be at least the length of the initial impulse response")
        return cls(h, lib)

    def update(self, response):
        r = _f32(response)
        _check(self._lib.lib.orc_crossfade_update(self._h, r, r.size), "response too long")

    def process(self, input, output):
        x = _f32(input)
        _check(self._lib.lib.orc_crossfade_process(self._h, x, x.size, output, output.size),
               "slice index out of range")

    def reset(self):
        _check(self._lib.lib.orc_crossfade_reset(self._h), "not yet implemented")

    def reset_ext(self):
        """EXTENSION beyond the reference (todo!() there): forget all audio, finish a running fade at once."""
        _check(self._lib.lib.orc_crossfade_reset_ext(self._h), "reset failed")

    def is_crossfading(self) -> bool:
        return bool(self._lib.lib.orc_crossfade_is_crossfading(self._h))

    @property
    def crossfader(self):
        return self._lib.lib.orc_crossfade_crossfader(self._h).contents

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.lib.orc_crossfade_free(self._h)
            self._h = None


def compute_tail_block_size(head_len: int, response_len: int) -> int:
    return load().lib.orc_compute_tail_block_size(head_len, response_len)


def gen_noise(channel: int, first_sample: int, n: int) -> np.ndarray:
    x = np.empty(n, dtype=np.float32)
    load().lib.orc_gen_noise(x, channel, first_sample, n)
    return x


def gen_ir(channel: int, update_index: int, length: int) -> np.ndarray:
    h = np.empty(length, dtype=np.float32)
    load().lib.orc_gen_ir(h, channel, update_index, length)
    return h


def direct_conv_f64(x, h) -> np.ndarray:
    x, h = _f32(x), _f32(h)
    y = np.empty(x.size, dtype=np.float64)
    load().lib.orc_direct_conv_f64(x, x.size, h, h.size, y)
    return y


def batch_twostage(irs, head_block: int, x, n_per_call: int, *, forced_tail: int = 0, irs_upd=None,
                   update_every: int = 0, threads: int | None = None) -> np.ndarray:
    """C independent TwoStageFFTConvolvers over x [C, calls*n_per_call] (OpenMP over channels)."""
    lib = load()
    irs, x = _f32(irs), _f32(x)
    Cn, L = irs.shape
    calls = x.shape[1] // n_per_call
    out = np.zeros((Cn, calls * n_per_call), np.float32)
    xin = np.ascontiguousarray(x[:, :calls * n_per_call])
    upd = _f32(irs_upd) if irs_upd is not None else None
    lib.lib.orc_batch_twostage_run(Cn, head_block, L, forced_tail, irs, upd.ctypes.data if upd is not None else None,
                                   upd.shape[0] if upd is not None else 0, update_every, xin, out, n_per_call, calls,
                                   threads or lib.lib.orc_max_threads())
    return out


def batch_crossfade(irs, block: int, x, *, irs_upd=None, update_every: int = 0, threads: int | None = None) -> np.ndarray:
    """C independent CrossfadeConvolver::init(h, block, len) over x [C, calls*block] (OpenMP over channels)."""
    lib = load()
    irs, x = _f32(irs), _f32(x)
    Cn, L = irs.shape
    calls = x.shape[1] // block
    out = np.zeros((Cn, calls * block), np.float32)
    xin = np.ascontiguousarray(x[:, :calls * block])
    upd = _f32(irs_upd) if irs_upd is not None else None
    lib.lib.orc_batch_crossfade_run(Cn, block, L, irs, upd.ctypes.data if upd is not None else None,
                                    upd.shape[0] if upd is not None else 0, update_every, xin, out, calls,
                                    threads or lib.lib.orc_max_threads())
    return out


def batch_fftconv(irs, block: int, x, n_per_call: int, threads: int | None = None) -> np.ndarray:
    """C independent FFTConvolvers over x [C, calls*n_per_call] (OpenMP over channels)."""
    lib = load()
    irs, x = _f32(irs), _f32(x)
    Cn, L = irs.shape
    calls = x.shape[1] // n_per_call
    out = np.zeros((Cn, calls * n_per_call), np.float32)
    lib.lib.orc_batch_fftconv_run(Cn, block, L, irs, np.ascontiguousarray(x[:, :calls * n_per_call]), out, n_per_call,
                                  calls, threads or lib.lib.orc_max_threads())
    return out
