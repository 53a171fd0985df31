"""fft_convolution_b200 — B200-native partitioned FFT convolution behind the `Convolution`
trait surface of Sin-tel/fft-convolution.  CUDA only (sm_100a); see DESIGN.md."""
from ._lib import ConvolutionPanic, CudaError, NotYetImplemented, load as load_library  # noqa: F401
from .convolvers import (  # noqa: F401
    CrossfadeConvolver, FFTConvolver, MimoConvolver, TwoStageFFTConvolver, compute_tail_block_size,
    mimo_segment_range, strict_todo, alloc_count,
)

__all__ = ["FFTConvolver", "TwoStageFFTConvolver", "CrossfadeConvolver", "MimoConvolver",
           "compute_tail_block_size", "mimo_segment_range", "strict_todo", "alloc_count",
           "ConvolutionPanic", "NotYetImplemented", "CudaError", "load_library"]
