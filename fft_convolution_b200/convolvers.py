"""Host-side mirror of the reference crate's `Convolution` trait (src/lib.rs:5-14) over the C ABI.

Same names and argument meaning as the reference:
    X.init(response, max_block_size, max_response_length) -> X
    x.update(response); x.reset(); x.process(input, output)
for X in FFTConvolver (src/fft_convolver.rs:86-307), TwoStageFFTConvolver (:323-526) and
CrossfadeConvolver (src/crossfade_convolver.rs:3-105); contract violations that panic in the
reference raise ConvolutionPanic.  The two methods the reference leaves `todo!()` (TwoStageFFTConvolver.update,
CrossfadeConvolver.reset) are implemented as documented extensions; `strict_todo(True)` makes them raise
NotYetImplemented like the reference.

Batched entry point: pass a 2-D response [channels, len]; input/output are then [channels, n]
(planar).  A 1-D response gives the reference's mono convolver.  All sample arithmetic runs in
the CUDA library; numpy arrays are only the host buffers handed across the boundary.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ConvolutionPanic, NotYetImplemented, Options, check  # noqa: F401


def _as_ir(response):
    r = np.ascontiguousarray(response, dtype=np.float32)
    if r.ndim == 1:
        return r.reshape(1, -1), True
    if r.ndim != 2:
        raise ValueError("response must be 1-D (mono) or 2-D [channels, len]")
    return r, False


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _opts(device=0, stream=None, shared_ir=False, async_tail=False, forced_tail_block=0, stages=0) -> Options:
    return Options(device, stream, int(shared_ir), int(async_tail), forced_tail_block, stages)


class _Base:
    _free = None

    def __init__(self, handle, channels: int, mono: bool):
        self._h = handle
        self.channels = channels
        self._mono = mono

    def _io(self, input, output):
        x = np.ascontiguousarray(input, dtype=np.float32)
        if not (isinstance(output, np.ndarray) and output.dtype == np.float32 and output.flags.c_contiguous):
            raise TypeError("output must be a C-contiguous float32 ndarray")
        if self._mono:
            if x.ndim != 1 or output.ndim != 1:
                raise ValueError("mono convolver expects 1-D input/output")
            return x, x.shape[0], x.shape[0], output, output.shape[0], output.shape[0]
        if x.ndim != 2 or output.ndim != 2 or x.shape[0] != self.channels or output.shape[0] != self.channels:
            raise ValueError(f"batched convolver expects [{self.channels}, n] input/output")
        return x, x.shape[1], x.shape[1], output, output.shape[1], output.shape[1]

    def _irs(self, response):
        r, _ = _as_ir(response)
        if r.shape[0] != self._ir_channels:
            raise ValueError(f"expected {self._ir_channels} impulse responses, got {r.shape[0]}")
        return r

    def close(self):
        if getattr(self, "_h", None):
            getattr(_lib.load(), self._free)(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FFTConvolver(_Base):
    """Uniformly partitioned convolver; reference: src/fft_convolver.rs:86-307."""
    _free = "fcb_fftconv_free"

    @classmethod
    def init(cls, response, block_size: int, max_response_length: int, *, channels: int | None = None,
             device: int = 0, stream=None) -> "FFTConvolver":
        """`channels` with a 1-D response = that many channels sharing one IR (stored once)."""
        lib = _lib.load()
        _lib.require_gpu()
        r, mono = _as_ir(response)
        shared = channels is not None and mono
        nch = channels if shared else r.shape[0]
        h = C.c_void_p()
        o = _opts(device, stream, shared_ir=shared)
        check(lib.fcb_fftconv_init(C.byref(h), _ptr(r), nch, r.shape[1], block_size, max_response_length, C.byref(o)))
        self = cls(h, nch, mono and not shared)
        self._ir_channels = 1 if shared else nch
        return self

    def clone(self) -> "FFTConvolver":
        h = C.c_void_p()
        check(_lib.load().fcb_fftconv_clone(self._h, C.byref(h)))
        c = FFTConvolver(h, self.channels, self._mono)
        c._ir_channels = self._ir_channels
        return c

    def update(self, response) -> None:
        r = self._irs(response)
        check(_lib.load().fcb_fftconv_update(self._h, _ptr(r), r.shape[1]))

    def reset(self) -> None:
        check(_lib.load().fcb_fftconv_reset(self._h))

    # real-time extras (no reference counterpart): process() and update() never allocate, so the workspace of
    # multi-block calls and the shadow spectra of background updates are reserved ahead of time
    def reserve(self, max_call_samples: int) -> None:
        """calls of up to this many samples may run as one time-batched pass"""
        check(_lib.load().fcb_fftconv_reserve(self._h, max_call_samples))

    def update_reserve(self) -> None:
        check(_lib.load().fcb_fftconv_update_reserve(self._h))

    def update_begin(self, response_ptr: int, ir_len: int, *, wait: bool = False) -> None:
        """background update from a raw pointer ([ir_channels][ir_len] f32, page-locked for a non-blocking call)"""
        check(_lib.load().fcb_fftconv_update_begin(self._h, response_ptr, ir_len, 1 if wait else 0))

    def update_pending(self) -> bool:
        return bool(_lib.load().fcb_fftconv_update_pending(self._h))

    def process(self, input, output) -> None:
        x, n_in, s_in, y, n_out, s_out = self._io(input, output)
        check(_lib.load().fcb_fftconv_process(self._h, _ptr(x), n_in, s_in, _ptr(y), n_out, s_out))

    def process_dev(self, in_ptr: int, in_len: int, in_stride: int, out_ptr: int, out_len: int, out_stride: int) -> None:
        """Asynchronous, device pointers (e.g. torch.Tensor.data_ptr())."""
        check(_lib.load().fcb_fftconv_process_dev(self._h, in_ptr, in_len, in_stride, out_ptr, out_len, out_stride, None))

    def sync(self) -> None:
        check(_lib.load().fcb_fftconv_sync(self._h))

    # scheduler scalars, for tests
    block_size = property(lambda s: _lib.load().fcb_fftconv_block_size(s._h))
    seg_count = property(lambda s: _lib.load().fcb_fftconv_seg_count(s._h))
    active_seg_count = property(lambda s: _lib.load().fcb_fftconv_active_seg_count(s._h))
    current = property(lambda s: _lib.load().fcb_fftconv_current(s._h))
    fill = property(lambda s: _lib.load().fcb_fftconv_fill(s._h))

    @property
    def engine(self):
        return _lib.load().fcb_fftconv_engine(self._h)

    def _read_row(self, fn, *idx):
        K = self.block_size + 1
        buf = np.empty(2 * K, np.float32)
        check(getattr(_lib.load(), fn)(self.engine, *idx, _ptr(buf)))
        return buf.view(np.complex64)

    def segment_ir(self, i, chan=0):
        return self._read_row("fcb_engine_read_ir_segment", chan, i)

    def segment(self, i, chan=0):
        return self._read_row("fcb_engine_read_ring_segment", chan, i)

    def premul(self, chan=0):
        return self._read_row("fcb_engine_read_premul", chan)

    def overlap(self, chan=0):
        buf = np.empty(self.block_size, np.float32)
        check(_lib.load().fcb_engine_read_overlap(self.engine, chan, _ptr(buf)))
        return buf

    def _release(self):
        h, self._h = self._h, None
        return h


class TwoStageFFTConvolver(_Base):
    """Two-stage (head/tail) partitioned convolver; reference: src/fft_convolver.rs:323-526."""
    _free = "fcb_twostage_free"

    @classmethod
    def init(cls, response, block_size: int, max_response_length: int, *, device: int = 0, stream=None,
             async_tail: bool = False, forced_tail_block: int = 0, stages: int = 2) -> "TwoStageFFTConvolver":
        """stages > 2 (extension): the partition nested — the tail is again a two-stage convolver (block sizes B, T1, T2, ...)"""
        lib = _lib.load()
        _lib.require_gpu()
        r, mono = _as_ir(response)
        h = C.c_void_p()
        o = _opts(device, stream, async_tail=async_tail, forced_tail_block=forced_tail_block, stages=stages)
        check(lib.fcb_twostage_init(C.byref(h), _ptr(r), r.shape[0], r.shape[1], block_size, max_response_length,
                                    C.byref(o)))
        self = cls(h, r.shape[0], mono)
        self._ir_channels = r.shape[0]
        return self

    def clone(self):
        h = C.c_void_p()
        check(_lib.load().fcb_twostage_clone(self._h, C.byref(h)))
        c = TwoStageFFTConvolver(h, self.channels, self._mono)
        c._ir_channels = self._ir_channels
        return c

    def update(self, response) -> None:
        r = self._irs(response)
        check(_lib.load().fcb_twostage_update(self._h, _ptr(r), r.shape[1]))

    def reset(self) -> None:
        check(_lib.load().fcb_twostage_reset(self._h))

    def process(self, input, output) -> None:
        x, n_in, s_in, y, n_out, s_out = self._io(input, output)
        check(_lib.load().fcb_twostage_process(self._h, _ptr(x), n_in, s_in, _ptr(y), n_out, s_out))

    def process_dev(self, in_ptr, in_len, in_stride, out_ptr, out_len, out_stride) -> None:
        check(_lib.load().fcb_twostage_process_dev(self._h, in_ptr, in_len, in_stride, out_ptr, out_len, out_stride))

    def sync(self) -> None:
        check(_lib.load().fcb_twostage_sync(self._h))

    @property
    def tail_block_size(self) -> int:
        return _lib.load().fcb_twostage_tail_block_size(self._h)

    @property
    def stage_blocks(self) -> list:
        buf = (C.c_size_t * 16)()
        n = _lib.load().fcb_twostage_stage_blocks(self._h, buf, 16)
        return [int(buf[i]) for i in range(min(n, 16))]


class CrossfadeConvolver(_Base):
    """Artefact-free IR switching over two FFTConvolvers; reference: src/crossfade_convolver.rs:3-105."""
    _free = "fcb_crossfade_free"

    @classmethod
    def new(cls, convolver: FFTConvolver, max_response_length: int, max_buffer_size: int,
            crossfade_samples: int) -> "CrossfadeConvolver":
        """Consumes `convolver` (the reference moves it in, src/crossfade_convolver.rs:20-42)."""
        h = C.c_void_p()
        channels, mono, irc = convolver.channels, convolver._mono, convolver._ir_channels
        check(_lib.load().fcb_crossfade_new(C.byref(h), convolver._h, max_response_length, max_buffer_size,
                                            crossfade_samples))
        convolver._release()
        self = cls(h, channels, mono)
        self._ir_channels = irc
        return self

    @classmethod
    def init(cls, response, max_block_size: int, max_response_length: int, *, device: int = 0, stream=None):
        lib = _lib.load()
        _lib.require_gpu()
        r, mono = _as_ir(response)
        h = C.c_void_p()
        o = _opts(device, stream)
        check(lib.fcb_crossfade_init(C.byref(h), _ptr(r), r.shape[0], r.shape[1], max_block_size,
                                     max_response_length, C.byref(o)))
        self = cls(h, r.shape[0], mono)
        self._ir_channels = r.shape[0]
        return self

    def update(self, response) -> None:
        r = self._irs(response)
        check(_lib.load().fcb_crossfade_update(self._h, _ptr(r), r.shape[1]))

    def reset(self) -> None:
        check(_lib.load().fcb_crossfade_reset(self._h))

    def process(self, input, output) -> None:
        x, n_in, s_in, y, n_out, s_out = self._io(input, output)
        check(_lib.load().fcb_crossfade_process(self._h, _ptr(x), n_in, s_in, _ptr(y), n_out, s_out))

    def process_dev(self, in_ptr, in_len, in_stride, out_ptr, out_len, out_stride) -> None:
        check(_lib.load().fcb_crossfade_process_dev(self._h, in_ptr, in_len, in_stride, out_ptr, out_len, out_stride))

    def sync(self) -> None:
        check(_lib.load().fcb_crossfade_sync(self._h))

    def is_crossfading(self) -> bool:
        return bool(_lib.load().fcb_crossfade_is_crossfading(self._h))

    def clone(self) -> "CrossfadeConvolver":
        h = C.c_void_p()
        check(_lib.load().fcb_crossfade_clone(self._h, C.byref(h)))
        c = CrossfadeConvolver(h, self.channels, self._mono)
        c._ir_channels = self._ir_channels
        return c

    def update_begin(self, response_ptr: int, ir_len: int) -> None:
        """update() that never waits: raw pointer to page-locked [ir_channels][ir_len] f32, kept alive by the caller"""
        check(_lib.load().fcb_crossfade_update_begin(self._h, response_ptr, ir_len))

    def update_pending(self) -> bool:
        return bool(_lib.load().fcb_crossfade_update_pending(self._h))

    def state(self):
        counter, mix, appr, tgt = C.c_int64(), C.c_float(), C.c_int(), C.c_int()
        check(_lib.load().fcb_crossfade_state(self._h, C.byref(counter), C.byref(mix), C.byref(appr), C.byref(tgt)))
        return counter.value, mix.value, bool(appr.value), tgt.value


def mimo_segment_range(seg_count: int, shard_index: int, shard_count: int) -> tuple[int, int]:
    """IR segments [lo, hi) owned by shard `shard_index` of `shard_count` (the C side's rule)."""
    return seg_count * shard_index // shard_count, seg_count * (shard_index + 1) // shard_count


class MimoConvolver:
    """OUT x IN convolution matrix: y_out = sum_in FFTConvolver(h[out][in]).process(x_in), built on
    the same K1/K2/K3 kernels with the IN input rings shared by all outputs.  `responses` is
    [OUT, IN, len].  With shard_count > 1 this object owns one contiguous range of IR segments and
    its partial spectra must be all-reduced across shards (see distributed.ShardedMimoConvolver)."""

    def __init__(self, handle, n_in, n_out, n_streams):
        self._h, self.n_in, self.n_out, self.n_streams = handle, n_in, n_out, n_streams

    @classmethod
    def init(cls, responses, block_size: int, max_response_length: int, *, n_streams: int = 1,
             shard_index: int = 0, shard_count: int = 1, device: int = 0, stream=None,
             tensor_cores: bool | None = None) -> "MimoConvolver":
        """tensor_cores: True / False force the tcgen05 matrix MAC (K4) on /
        off for this object; None keeps the library default (on from 33 streams; fewer run k_mac_rt on the FP32 pipes)."""
        lib = _lib.load()
        _lib.require_gpu()
        r = np.ascontiguousarray(responses, dtype=np.float32)
        if r.ndim != 3:
            raise ValueError("responses must be [out, in, len]")
        n_out, n_in, length = r.shape
        d = _lib.MimoDesc(n_in, n_out, n_streams, block_size, max_response_length, shard_index, shard_count,
                          device, stream)
        h = C.c_void_p()
        if tensor_cores is not None:
            check(lib.fcb_tune(b"mimo_tc", 1 if tensor_cores else 0))
        try:
            check(lib.fcb_mimo_create(C.byref(d), C.byref(h)))
        finally:
            if tensor_cores is not None:
                check(lib.fcb_tune(b"mimo_tc", 2))
        self = cls(h, n_in, n_out, n_streams)
        check(lib.fcb_mimo_set_ir(h, _ptr(r), length))
        check(lib.fcb_mimo_sync(h))
        return self

    block_size = property(lambda s: _lib.load().fcb_mimo_block_size(s._h))
    seg_count = property(lambda s: _lib.load().fcb_mimo_seg_count(s._h))
    uses_tensor_cores = property(lambda s: _lib.load().fcb_mimo_uses_tensor_cores(s._h) == 1)
    #: which delay-line MAC this object runs: "tensor" (K4, tcgen05), "register_tile" (k_mac_rt) or "tile" (k_mac_tile / K2)
    mac_kernel = property(lambda s: {1: "tensor", 2: "register_tile"}.get(_lib.load().fcb_mimo_uses_tensor_cores(s._h), "tile"))

    @property
    def segment_range(self):
        lo, hi = C.c_size_t(), C.c_size_t()
        check(_lib.load().fcb_mimo_segment_range(self._h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def process(self, input, output) -> None:
        """input [NS*IN, B] -> output [NS*OUT, B] (host arrays, one full block)."""
        x = np.ascontiguousarray(input, dtype=np.float32)
        B = self.block_size
        if x.shape != (self.n_streams * self.n_in, B) or output.shape != (self.n_streams * self.n_out, B):
            raise ValueError("MimoConvolver.process expects one full block: [NS*IN, B] -> [NS*OUT, B]")
        if not (output.dtype == np.float32 and output.flags.c_contiguous):
            raise TypeError("output must be a C-contiguous float32 ndarray")
        check(_lib.load().fcb_mimo_process(self._h, _ptr(x), _ptr(output)))

    def reset(self) -> None:
        check(_lib.load().fcb_mimo_reset(self._h))

    def set_ir(self, responses) -> None:
        """replace the whole matrix ([OUT, IN, len], len <= max_response_length); input history is kept like update()"""
        r = np.ascontiguousarray(responses, dtype=np.float32)
        if r.ndim != 3 or r.shape[:2] != (self.n_out, self.n_in):
            raise ValueError("responses must be [out, in, len]")
        check(_lib.load().fcb_mimo_set_ir(self._h, _ptr(r), r.shape[2]))
        check(_lib.load().fcb_mimo_sync(self._h))

    def partial_dev(self, in_ptr: int, in_stride: int) -> None:
        check(_lib.load().fcb_mimo_partial_dev(self._h, in_ptr, in_stride))

    def conv_buffer(self) -> tuple[int, int]:
        n = C.c_size_t()
        p = _lib.load().fcb_mimo_conv_buffer(self._h, C.byref(n))
        return p, n.value

    def finish_dev(self, out_ptr: int, out_stride: int) -> None:
        check(_lib.load().fcb_mimo_finish_dev(self._h, out_ptr, out_stride))

    def sync(self) -> None:
        check(_lib.load().fcb_mimo_sync(self._h))

    # peer exchange of the partial spectra (IR-partition shards on one NVLink node)
    def peer_export(self) -> bytes:
        buf = (C.c_ubyte * 64)()
        check(_lib.load().fcb_mimo_peer_export(self._h, buf))
        return bytes(buf)

    def peer_attach(self, handles) -> None:
        blob = b"".join(handles)
        check(_lib.load().fcb_mimo_peer_attach(self._h, C.cast(C.c_char_p(blob), C.c_void_p)))

    def peer_inbox(self) -> int:
        return _lib.load().fcb_mimo_peer_inbox(self._h)

    def peer_set_scatter(self, on: bool) -> None:
        """reduce-scatter form of the peer exchange: this shard finishes (and outputs) only its own rows"""
        check(_lib.load().fcb_mimo_peer_set_scatter(self._h, 1 if on else 0))

    def set_overlap(self, on: bool) -> None:
        """peer exchange: K3 (the kernel that waits for the peers) on its own stream beside the next block's MAC;
        outputs are ordered on the convolver's stream by join(), for the host by sync()"""
        check(_lib.load().fcb_mimo_set_overlap(self._h, 1 if on else 0))

    def join(self) -> None:
        check(_lib.load().fcb_mimo_join(self._h))

    @property
    def owned_rows(self):
        lo, hi = C.c_size_t(), C.c_size_t()
        check(_lib.load().fcb_mimo_owned_rows(self._h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def finish_rows_dev(self, out_ptr: int, out_stride: int, row_lo: int, row_hi: int) -> None:
        check(_lib.load().fcb_mimo_finish_rows_dev(self._h, out_ptr, out_stride, row_lo, row_hi))

    def peer_attach_ptrs(self, inboxes) -> None:
        arr = (C.c_void_p * len(inboxes))(*inboxes)
        check(_lib.load().fcb_mimo_peer_attach_ptrs(self._h, arr))

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().fcb_mimo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def compute_tail_block_size(head_len: int, response_len: int) -> int:
    """src/fft_convolver.rs:520-526 (f32 arithmetic)."""
    return _lib.load().fcb_compute_tail_block_size(head_len, response_len)


def strict_todo(on: bool) -> None:
    """True: TwoStageFFTConvolver.update / CrossfadeConvolver.reset raise NotYetImplemented like the reference's todo!()."""
    check(_lib.load().fcb_tune(b"strict_todo", 1 if on else 0))


def alloc_count() -> int:
    """allocations / stream and event creations / releases made by the library so far (real-time tests)"""
    return _lib.load().fcb_debug_alloc_count()
