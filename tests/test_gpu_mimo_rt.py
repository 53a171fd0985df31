"""k_mac_rt (mimo_rt.cuh) — the convolution-matrix MAC for a few streams as a register-tiled per-bin complex GEMM on the
FP32 pipes — against OUT x IN oracle convolvers per stream, against the shared-memory tile kernel it replaces, sharded by
IR partition, and through reset / a second set_ir.  Tolerance is north_star's: max |err| <= 1e-5 x output RMS."""
import numpy as np
import pytest

import oracle
from mimo_oracle import MimoOracle
from refsignals import rms

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def F():
    import fft_convolution_b200 as f
    return f


def _irs(n_out, n_in, L, upd=0):
    return np.stack([np.stack([oracle.gen_ir(o * n_in + i, upd, L) for i in range(n_in)]) for o in range(n_out)])


def _run(g, x, B, nblocks, n_rows_out):
    y = np.zeros((n_rows_out, B * nblocks), np.float32)
    blk_out = np.zeros((n_rows_out, B), np.float32)
    for b in range(nblocks):
        g.process(np.ascontiguousarray(x[:, b * B:(b + 1) * B]), blk_out)
        y[:, b * B:(b + 1) * B] = blk_out
    return y


def _worst(y, ref):
    return float(np.max(np.max(np.abs(y - ref), axis=1) / np.array([rms(r) for r in ref])))


# every CTA shape (8 / 16 outputs x 4 / 8 / 16 streams), ragged output and stream counts, the packed bin-0 lane, one and
# several bin tiles, delay lines shorter and longer than a stage, nblocks > S so `current` takes every value (the ring
# wrap falls into every position of a box)
@pytest.mark.parametrize("n_out,n_in,B,L,NS,nblocks", [
    (16, 3, 64, 64 * 25 + 1, 1, 29),     # one stream (BASELINE configs[4] as worded): <2,1>
    (3, 2, 64, 64 * 13 + 5, 2, 20),      # <1,1>
    (8, 3, 32, 32 * 9, 5, 14),           # <1,2>
    (5, 1, 128, 128 * 6 + 1, 11, 10),    # <1,4>
    (9, 2, 64, 64 * 21 + 3, 3, 26),      # <2,1>
    (16, 2, 64, 64 * 10, 8, 13),         # <2,2>
    (16, 3, 32, 32 * 37 + 5, 16, 41),    # <2,4>: the headline shape of the kernel in small
    (20, 2, 32, 32 * 5, 19, 8),          # two output groups, two stream groups
    (2, 2, 1024, 1024 * 3 + 7, 2, 6),    # 32 bin tiles, two segments past segment 0
])
def test_rt_matches_oracle(F, n_out, n_in, B, L, NS, nblocks):
    h = _irs(n_out, n_in, L)
    x = np.stack([oracle.gen_noise(900 + i, 0, B * nblocks) for i in range(NS * n_in)])
    g = F.MimoConvolver.init(h, B, L, n_streams=NS, tensor_cores=False)
    assert g.mac_kernel == "register_tile"
    y = _run(g, x, B, nblocks, NS * n_out)
    worst = 0.0
    for s in sorted({0, NS // 2, NS - 1}):
        ref = MimoOracle(h, B, L).process(x[s * n_in:(s + 1) * n_in])
        worst = max(worst, _worst(y[s * n_out:(s + 1) * n_out], ref))
    print(f"register-tiled matrix MAC vs oracle: max |err| / rms = {worst:.3e}")
    assert worst <= TOL


def test_rt_equals_the_tile_kernel_and_the_tensor_cores(F):
    """all three matrix MACs on one problem, every stream and output row compared"""
    from fft_convolution_b200 import _lib
    lib = _lib.load()
    n_out, n_in, B, L, NS, nblocks = 16, 4, 256, 256 * 21 + 3, 7, 30
    h = _irs(n_out, n_in, L)
    x = np.stack([oracle.gen_noise(600 + i, 0, B * nblocks) for i in range(NS * n_in)])
    ys = {}
    try:
        for kind in ("register_tile", "tile", "tensor"):
            _lib.check(lib.fcb_tune(b"mimo_rt", 0 if kind == "tile" else 1))
            g = F.MimoConvolver.init(h, B, L, n_streams=NS, tensor_cores=kind == "tensor")
            assert g.mac_kernel == kind
            ys[kind] = _run(g, x, B, nblocks, NS * n_out)
    finally:
        _lib.check(lib.fcb_tune(b"mimo_rt", 1))
    assert _worst(ys["register_tile"], ys["tile"]) <= TOL
    assert _worst(ys["register_tile"], ys["tensor"]) <= TOL


def test_rt_streams_are_independent_and_symmetric(F):
    """the same input on every stream gives the same bits on every stream (each stream's sum is formed in the same order)"""
    n_out, n_in, B, L, NS, nblocks = 11, 2, 64, 64 * 17 + 9, 6, 22
    h = _irs(n_out, n_in, L)
    one = np.stack([oracle.gen_noise(950 + i, 0, B * nblocks) for i in range(n_in)])
    x = np.concatenate([one] * NS)
    g = F.MimoConvolver.init(h, B, L, n_streams=NS, tensor_cores=False)
    assert g.mac_kernel == "register_tile"
    y = _run(g, x, B, nblocks, NS * n_out)
    for s in range(1, NS):
        assert np.array_equal(y[:n_out], y[s * n_out:(s + 1) * n_out])
    assert _worst(y[:n_out], MimoOracle(h, B, L).process(one)) <= TOL


def test_rt_reset_and_set_ir_again(F):
    n_out, n_in, B, L, NS = 16, 2, 64, 64 * 9, 4
    h = _irs(n_out, n_in, L)
    x = np.stack([oracle.gen_noise(700 + i, 0, B * 12) for i in range(NS * n_in)])
    g = F.MimoConvolver.init(h, B, L, n_streams=NS, tensor_cores=False)
    assert g.mac_kernel == "register_tile"
    a = _run(g, x, B, 12, NS * n_out)
    g.reset()
    b = _run(g, x, B, 12, NS * n_out)
    assert np.array_equal(a, b)
    h2 = _irs(n_out, n_in, L, upd=1)
    g.set_ir(h2)
    g.reset()
    c = _run(g, x, B, 12, NS * n_out)
    assert _worst(c[:n_out], MimoOracle(h2, B, L).process(x[:n_in])) <= TOL


@pytest.mark.parametrize("shards", [2, 3, 8])
def test_rt_ir_partition_shards(F, shards):
    """every shard runs k_mac_rt over its own segment range (the first one without segment 0, which belongs to the reduce
    kernel); partial spectra summed as the exchange would, K3 once — equal to the unsharded engine to f32 rounding"""
    import torch
    from fft_convolution_b200.distributed import _DeviceBuffer
    n_out, n_in, B, L, NS, nblocks = 9, 2, 64, 64 * 29 + 5, 5, 34
    h = _irs(n_out, n_in, L)
    x = np.stack([oracle.gen_noise(800 + i, 0, B * nblocks) for i in range(NS * n_in)])
    whole = F.MimoConvolver.init(h, B, L, n_streams=NS, tensor_cores=False)
    parts = [F.MimoConvolver.init(h, B, L, n_streams=NS, shard_index=g, shard_count=shards, tensor_cores=False) for g in range(shards)]
    assert whole.mac_kernel == "register_tile" and all(p.mac_kernel == "register_tile" for p in parts)
    yw = _run(whole, x, B, nblocks, NS * n_out)
    d_in = torch.empty((NS * n_in, B), dtype=torch.float32, device="cuda")
    d_out = torch.empty((NS * n_out, B), dtype=torch.float32, device="cuda")
    ys = np.zeros_like(yw)
    for b in range(nblocks):
        d_in.copy_(torch.from_numpy(np.ascontiguousarray(x[:, b * B:(b + 1) * B])))
        torch.cuda.synchronize()
        bufs = []
        for p in parts:
            p.partial_dev(d_in.data_ptr(), B)
            p.sync()
            ptr, n = p.conv_buffer()
            bufs.append(torch.as_tensor(_DeviceBuffer(ptr, n), device="cuda"))
        total = torch.stack(bufs).sum(dim=0)  # the all-reduce
        for p, bview in zip(parts, bufs):
            bview.copy_(total)
        torch.cuda.synchronize()
        for p in parts:
            p.finish_dev(d_out.data_ptr(), B)
            p.sync()
        ys[:, b * B:(b + 1) * B] = d_out.cpu().numpy()
    assert _worst(ys, yw) <= TOL
    assert _worst(ys[:n_out], MimoOracle(h, B, L).process(x[:n_in])) <= TOL
