"""CPU oracle for the partitioned-FFT-convolution hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (fft_convolution_b200) never does.
"""
from .oracle_c import (  # noqa: F401
    OracleLib, load, build,
    FFTConvolver, TwoStageFFTConvolver, CrossfadeConvolver, Crossfader,
    compute_tail_block_size, gen_noise, gen_ir, direct_conv_f64, OraclePanic,
    batch_twostage, batch_crossfade, batch_fftconv,
)
