"""K4 — the convolution-matrix MAC as per-bin complex GEMMs on the tensor cores (tcgen05, 3xTF32,
mimo_tc.cuh) — against OUT x IN oracle convolvers per stream, against the CUDA-core matrix kernel,
and sharded by IR partition.  Tolerance is north_star's: max |err| <= 1e-5 x output RMS."""
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


@pytest.mark.parametrize("n_in,B,L,NS,nblocks", [(3, 64, 64 * 37 + 5, 5, 45), (1, 32, 32 * 16, 2, 20), (2, 128, 128 * 3 + 1, 3, 9)])
def test_tc_matches_oracle(F, n_in, B, L, NS, nblocks):
    """every stream of the tensor-core engine == 16 x IN reference convolvers; nblocks > S so the
    ring wraps and `current` takes every value (both slot ranges, ragged 16-segment chunks)"""
    n_out = 16
    h = _irs(n_out, n_in, L)
    x = np.stack([oracle.gen_noise(500 + i, 0, B * nblocks) for i in range(NS * n_in)])
    g = F.MimoConvolver.init(h, B, L, n_streams=NS, tensor_cores=True)
    assert g.uses_tensor_cores
    y = _run(g, x, B, nblocks, NS * n_out)
    worst = 0.0
    for s in range(NS):
        ref = MimoOracle(h, B, L).process(x[s * n_in:(s + 1) * n_in])
        err = np.max(np.abs(y[s * n_out:(s + 1) * n_out] - ref), axis=1) / np.array([rms(r) for r in ref])
        worst = max(worst, float(err.max()))
    print(f"tensor-core matrix MAC vs oracle: max |err| / rms = {worst:.3e}")
    assert worst <= TOL


@pytest.mark.parametrize("n_out,n_in,B,L,NS,nblocks", [(5, 2, 64, 64 * 20 + 3, 3, 24),      # outputs padded to one group of 16
                                                       (20, 2, 32, 32 * 18, 2, 21),        # two output groups
                                                       (16, 1, 32, 32 * 17 + 1, 131, 19)])  # two stream groups of 128
def test_tc_any_output_and_stream_count(F, n_out, n_in, B, L, NS, nblocks):
    h = _irs(n_out, n_in, L)
    rng = np.random.default_rng(NS)
    x = (rng.random((NS * n_in, B * nblocks), dtype=np.float32) * 2 - 1).astype(np.float32)
    g = F.MimoConvolver.init(h, B, L, n_streams=NS, tensor_cores=True)
    assert g.uses_tensor_cores
    y = _run(g, x, B, nblocks, NS * n_out)
    worst = 0.0
    for s in sorted({0, 1, NS // 2, NS - 1}):
        ref = MimoOracle(h, B, L).process(x[s * n_in:(s + 1) * n_in])
        err = np.max(np.abs(y[s * n_out:(s + 1) * n_out] - ref), axis=1) / np.array([rms(r) for r in ref])
        worst = max(worst, float(err.max()))
    assert worst <= TOL


def test_tc_equals_cuda_core_matrix_kernel(F):
    n_out, n_in, B, L, NS, nblocks = 16, 4, 256, 256 * 21 + 3, 7, 30
    h = _irs(n_out, n_in, L)
    x = np.stack([oracle.gen_noise(600 + i, 0, B * nblocks) for i in range(NS * n_in)])
    ys = {}
    for tc in (True, False):
        g = F.MimoConvolver.init(h, B, L, n_streams=NS, tensor_cores=tc)
        assert g.uses_tensor_cores == tc
        ys[tc] = _run(g, x, B, nblocks, NS * n_out)
    r = np.array([rms(v) for v in ys[False]])
    assert np.max(np.max(np.abs(ys[True] - ys[False]), axis=1) / r) <= TOL


def test_tc_reset_and_set_ir_again(F):
    n_out, n_in, B, L, NS = 16, 2, 64, 64 * 9, 2
    h = _irs(n_out, n_in, L)
    x = np.stack([oracle.gen_noise(700 + i, 0, B * 12) for i in range(NS * n_in)])
    g = F.MimoConvolver.init(h, B, L, n_streams=NS, tensor_cores=True)
    a = _run(g, x, B, 12, NS * n_out)
    g.reset()
    b = _run(g, x, B, 12, NS * n_out)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("shards", [2, 3])
def test_tc_ir_partition_shards(F, shards):
    """tensor-core shards own S*g/G..S*(g+1)/G of the segments; summed partial spectra (the NCCL
    all-reduce) then K3 == the oracle"""
    import torch
    from fft_convolution_b200.distributed import _DeviceBuffer
    n_out, n_in, B, L = 16, 2, 64, 64 * 41 + 7
    h = _irs(n_out, n_in, L)
    nblocks = 50
    x = np.stack([oracle.gen_noise(800 + i, 0, B * nblocks) for i in range(n_in)])
    parts = [F.MimoConvolver.init(h, B, L, shard_index=g, shard_count=shards, tensor_cores=True) for g in range(shards)]
    assert all(p.uses_tensor_cores for p in parts)
    ref = MimoOracle(h, B, L).process(x)
    d_in = torch.empty((n_in, B), dtype=torch.float32, device="cuda")
    d_out = torch.empty((n_out, B), dtype=torch.float32, device="cuda")
    y = np.zeros_like(ref)
    for b in range(nblocks):
        d_in.copy_(torch.from_numpy(np.ascontiguousarray(x[:, b * B:(b + 1) * B])))
        torch.cuda.synchronize()
        bufs = []
        for p in parts:
            p.partial_dev(d_in.data_ptr(), B)
            p.sync()
            ptr, n = p.conv_buffer()
            bufs.append(torch.as_tensor(_DeviceBuffer(ptr, n), device="cuda"))
        total = torch.stack(bufs).sum(dim=0)
        for bview in bufs:
            bview.copy_(total)
        torch.cuda.synchronize()
        for p in parts:
            p.finish_dev(d_out.data_ptr(), B)
            p.sync()
        y[:, b * B:(b + 1) * B] = d_out.cpu().numpy()
    r = np.array([rms(v) for v in ref])
    assert np.max(np.max(np.abs(y - ref), axis=1) / r) <= TOL


def test_tc_128_streams_block_512(F):
    """the M = 128 tile full: 128 streams x 16 x 4 matrix at block 512 vs the CUDA-core kernel"""
    n_out, n_in, B, L, NS, nblocks = 16, 4, 512, 512 * 18, 128, 4
    h = _irs(n_out, n_in, L)
    rng = np.random.default_rng(7)
    x = (rng.random((NS * n_in, B * nblocks), dtype=np.float32) * 2 - 1).astype(np.float32)
    ys = {}
    for tc in (True, False):
        g = F.MimoConvolver.init(h, B, L, n_streams=NS, tensor_cores=tc)
        ys[tc] = _run(g, x, B, nblocks, NS * n_out)
    r = np.array([max(rms(v), 1e-3) for v in ys[False]])
    assert np.max(np.max(np.abs(ys[True] - ys[False]), axis=1) / r) <= TOL
