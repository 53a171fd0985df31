"""ctypes binding of libfftconv_b200.so (the C ABI declared in include/fftconv_b200.h).

No fallback: if the shared library is missing or no CUDA device is visible, using the package
raises.  The library is built in-tree by fft_convolution_b200/build.py (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
SO = PKG / "libfftconv_b200.so"

FCB_OK, FCB_ERR_PANIC, FCB_ERR_TODO, FCB_ERR_CUDA, FCB_ERR_UNSUPPORTED, FCB_ERR_ARG = range(6)


class ConvolutionPanic(RuntimeError):
    """Contract violation: the reference implementation panics here."""


class NotYetImplemented(ConvolutionPanic):
    """The reference method is `todo!()`."""


class CudaError(RuntimeError):
    pass


class Options(C.Structure):
    _fields_ = [("device", C.c_int), ("stream", C.c_void_p), ("shared_ir", C.c_int),
                ("async_tail", C.c_int), ("forced_tail_block", C.c_size_t), ("stages", C.c_size_t)]


class EngineDesc(C.Structure):
    _fields_ = [("channels", C.c_size_t), ("block_size", C.c_size_t), ("max_response_length", C.c_size_t),
                ("shared_ir", C.c_int), ("device", C.c_int), ("stream", C.c_void_p)]


class MimoDesc(C.Structure):
    _fields_ = [("n_in", C.c_size_t), ("n_out", C.c_size_t), ("n_streams", C.c_size_t),
                ("block_size", C.c_size_t), ("max_response_length", C.c_size_t),
                ("shard_index", C.c_size_t), ("shard_count", C.c_size_t), ("device", C.c_int), ("stream", C.c_void_p)]


class Epilogue(C.Structure):
    _fields_ = [("add0", C.c_void_p), ("add1", C.c_void_p), ("add_stride", C.c_size_t),
                ("mix_other", C.c_void_p), ("mix_stride", C.c_size_t), ("gains", C.c_void_p)]


_sz, _vp, _i = C.c_size_t, C.c_void_p, C.c_int
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol include/fftconv_b200.h declares
SIGNATURES = {
    "fcb_last_error": (C.c_char_p, []),
    "fcb_version": (C.c_char_p, []),
    "fcb_launch_count": (C.c_uint64, []),
    "fcb_device_count": (_i, []),
    "fcb_debug_alloc_count": (C.c_uint64, []),
    "fcb_tune": (_i, [C.c_char_p, _i]),
    "fcb_profile_mac": (_i, [_i]),
    "fcb_profile_mac_read": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "fcb_host_alloc": (_vp, [_sz]),
    "fcb_host_free": (None, [_vp]),
    "fcb_engine_create": (_i, [C.POINTER(EngineDesc), _pp]),
    "fcb_engine_destroy": (None, [_vp]),
    "fcb_engine_clone": (_i, [_vp, _pp]),
    "fcb_engine_set_stream": (_i, [_vp, _vp]),
    "fcb_engine_stream": (_vp, [_vp]),
    "fcb_engine_sync": (_i, [_vp]),
    "fcb_engine_channels": (_sz, [_vp]),
    "fcb_engine_block_size": (_sz, [_vp]),
    "fcb_engine_seg_count": (_sz, [_vp]),
    "fcb_engine_set_ir": (_i, [_vp, _sz, _sz, _vp, _sz, _sz, _i]),
    "fcb_engine_set_ir_dev": (_i, [_vp, _sz, _sz, _vp, _sz, _sz, _i]),
    "fcb_engine_update_reserve": (_i, [_vp]),
    "fcb_engine_update_reserved": (_i, [_vp]),
    "fcb_engine_update_begin": (_i, [_vp, _vp, _sz, _sz, _i]),
    "fcb_engine_update_ready": (_i, [_vp]),
    "fcb_engine_update_commit": (_i, [_vp]),
    "fcb_engine_update_wait": (_i, [_vp]),
    "fcb_engine_update_join": (_i, [_vp, _vp]),
    "fcb_engine_reset": (_i, [_vp]),
    "fcb_engine_push_input": (_i, [_vp, _vp, _sz, _sz, _sz]),
    "fcb_engine_push_input_dev": (_i, [_vp, _vp, _sz, _sz, _sz]),
    "fcb_engine_fft_forward": (_i, [_vp, _sz, _sz]),
    "fcb_engine_mac": (_i, [_vp, _sz, _sz]),
    "fcb_engine_ifft_ola": (_i, [_vp, _sz, _sz, _sz, _i, _vp, _sz, C.POINTER(Epilogue)]),
    "fcb_engine_scratch": (_vp, [_vp]),
    "fcb_engine_input_buffer": (_vp, [_vp]),
    "fcb_engine_fetch": (_i, [_vp, _vp, _sz, _vp, _sz, _sz]),
    "fcb_engine_process_block_dev": (_i, [_vp, _vp, _sz, _vp, _sz, _sz, _sz, C.POINTER(Epilogue)]),
    "fcb_engine_process_block_host": (_i, [_vp, _vp, _sz, _vp, _sz, _sz, _sz, _sz]),
    "fcb_engine_pair_ok": (_i, [_vp, _vp, _sz]),
    "fcb_engine_process_block_pair_dev": (_i, [_vp, _vp, _vp, _sz, _vp, _sz, C.POINTER(Epilogue), _vp, _sz, C.POINTER(Epilogue), _sz, _sz]),
    "fcb_engine_process_block_pair_copy_dev": (_i, [_vp, _vp, _vp, _sz, _vp, _sz, C.POINTER(Epilogue), _vp, _sz, C.POINTER(Epilogue), _sz, _sz, _vp, _sz]),
    "fcb_engine_multi_block_ok": (_i, [_vp, _sz, _sz]),
    "fcb_engine_multi_block_capacity": (_sz, [_vp]),
    "fcb_engine_multi_block_reserved": (_sz, [_vp]),
    "fcb_engine_multi_block_reserve": (_i, [_vp, _sz]),
    "fcb_engine_process_blocks": (_i, [_vp, _vp, _sz, _vp, _sz, _sz, _sz, _sz, C.POINTER(Epilogue), _i]),
    "fcb_engine_read_ir_segment": (_i, [_vp, _sz, _sz, _vp]),
    "fcb_engine_read_ring_segment": (_i, [_vp, _sz, _sz, _vp]),
    "fcb_engine_read_premul": (_i, [_vp, _sz, _vp]),
    "fcb_engine_read_overlap": (_i, [_vp, _sz, _vp]),
    "fcb_engine_write_ir_segment": (_i, [_vp, _sz, _sz, _vp]),
    "fcb_engine_write_ring_segment": (_i, [_vp, _sz, _sz, _vp]),
    "fcb_fftconv_init": (_i, [_pp, _vp, _sz, _sz, _sz, _sz, C.POINTER(Options)]),
    "fcb_fftconv_default": (_i, [_pp, _sz, C.POINTER(Options)]),
    "fcb_fftconv_clone": (_i, [_vp, _pp]),
    "fcb_fftconv_free": (None, [_vp]),
    "fcb_fftconv_update": (_i, [_vp, _vp, _sz]),
    "fcb_fftconv_update_reserve": (_i, [_vp]),
    "fcb_fftconv_update_begin": (_i, [_vp, _vp, _sz, _i]),
    "fcb_fftconv_update_pending": (_i, [_vp]),
    "fcb_fftconv_reserve": (_i, [_vp, _sz]),
    "fcb_fftconv_reset": (_i, [_vp]),
    "fcb_fftconv_process": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz]),
    "fcb_fftconv_process_dev": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, C.POINTER(Epilogue)]),
    "fcb_fftconv_sync": (_i, [_vp]),
    "fcb_fftconv_engine": (_vp, [_vp]),
    "fcb_fftconv_block_size": (_sz, [_vp]),
    "fcb_fftconv_seg_count": (_sz, [_vp]),
    "fcb_fftconv_active_seg_count": (_sz, [_vp]),
    "fcb_fftconv_current": (_sz, [_vp]),
    "fcb_fftconv_fill": (_sz, [_vp]),
    "fcb_compute_tail_block_size": (_sz, [_sz, _sz]),
    "fcb_twostage_init": (_i, [_pp, _vp, _sz, _sz, _sz, _sz, C.POINTER(Options)]),
    "fcb_twostage_clone": (_i, [_vp, _pp]),
    "fcb_twostage_free": (None, [_vp]),
    "fcb_twostage_update": (_i, [_vp, _vp, _sz]),
    "fcb_twostage_reset": (_i, [_vp]),
    "fcb_twostage_process": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz]),
    "fcb_twostage_process_dev": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz]),
    "fcb_twostage_sync": (_i, [_vp]),
    "fcb_twostage_tail_block_size": (_sz, [_vp]),
    "fcb_twostage_stage_blocks": (_sz, [_vp, C.POINTER(_sz), _sz]),
    "fcb_crossfade_new": (_i, [_pp, _vp, _sz, _sz, _sz]),
    "fcb_crossfade_init": (_i, [_pp, _vp, _sz, _sz, _sz, _sz, C.POINTER(Options)]),
    "fcb_crossfade_free": (None, [_vp]),
    "fcb_crossfade_update": (_i, [_vp, _vp, _sz]),
    "fcb_crossfade_process": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz]),
    "fcb_crossfade_process_dev": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz]),
    "fcb_crossfade_reset": (_i, [_vp]),
    "fcb_crossfade_clone": (_i, [_vp, _pp]),
    "fcb_crossfade_update_begin": (_i, [_vp, _vp, _sz]),
    "fcb_crossfade_update_pending": (_i, [_vp]),
    "fcb_crossfade_is_crossfading": (_i, [_vp]),
    "fcb_crossfade_sync": (_i, [_vp]),
    "fcb_crossfade_state": (_i, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(_i), C.POINTER(_i)]),
    "fcb_mimo_create": (_i, [C.POINTER(MimoDesc), _pp]),
    "fcb_mimo_destroy": (None, [_vp]),
    "fcb_mimo_set_ir": (_i, [_vp, _vp, _sz]),
    "fcb_mimo_reset": (_i, [_vp]),
    "fcb_mimo_partial_dev": (_i, [_vp, _vp, _sz]),
    "fcb_mimo_conv_buffer": (_vp, [_vp, C.POINTER(_sz)]),
    "fcb_mimo_finish_dev": (_i, [_vp, _vp, _sz]),
    "fcb_mimo_process": (_i, [_vp, _vp, _vp]),
    "fcb_mimo_sync": (_i, [_vp]),
    "fcb_mimo_stream": (_vp, [_vp]),
    "fcb_debug_tc_stages": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "fcb_debug_mac_tile_plan": (_i, [_i, _i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "fcb_mimo_uses_tensor_cores": (_i, [_vp]),
    "fcb_mimo_peer_export": (_i, [_vp, _vp]),
    "fcb_mimo_peer_attach": (_i, [_vp, _vp]),
    "fcb_mimo_peer_inbox": (_vp, [_vp]),
    "fcb_mimo_peer_set_scatter": (_i, [_vp, _i]),
    "fcb_mimo_set_overlap": (_i, [_vp, _i]),
    "fcb_mimo_join": (_i, [_vp]),
    "fcb_mimo_owned_rows": (_i, [_vp, C.POINTER(_sz), C.POINTER(_sz)]),
    "fcb_mimo_finish_rows_dev": (_i, [_vp, _vp, _sz, _sz, _sz]),
    "fcb_mimo_peer_attach_ptrs": (_i, [_vp, _vp]),
    "fcb_mimo_block_size": (_sz, [_vp]),
    "fcb_mimo_seg_count": (_sz, [_vp]),
    "fcb_mimo_segment_range": (_i, [_vp, C.POINTER(_sz), C.POINTER(_sz)]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library and bind every exported symbol; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not SO.exists():
        raise ImportError(
            f"{SO} is missing: build it with `python -m fft_convolution_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(str(SO))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc == FCB_OK:
        return
    msg = load().fcb_last_error().decode(errors="replace")
    if rc == FCB_ERR_PANIC:
        raise ConvolutionPanic(msg)
    if rc == FCB_ERR_TODO:
        raise NotYetImplemented(msg)
    if rc == FCB_ERR_CUDA:
        raise CudaError(msg)
    if rc == FCB_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise ValueError(msg)


def require_gpu() -> None:
    n = load().fcb_device_count()
    if n <= 0:
        raise CudaError("no CUDA device visible: fft_convolution_b200 has no CPU path")
