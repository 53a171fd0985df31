"""IR-partition sharding of the convolution matrix, CPU side: the segment-range rule partitions
[0, S), and per-rank partial spectra summed with a gloo all-reduce (world size 2) equal the
unsharded delay-line sum (src/fft_convolver.rs:244-261 is associative over segments)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fft_convolution_b200.convolvers import mimo_segment_range

S, K, N_IN, N_OUT = 11, 33, 2, 3


def _spectra(seed):
    rng = np.random.default_rng(seed)
    H = (rng.standard_normal((N_OUT, N_IN, S, K)) + 1j * rng.standard_normal((N_OUT, N_IN, S, K))).astype(np.complex64)
    X = (rng.standard_normal((N_IN, S, K)) + 1j * rng.standard_normal((N_IN, S, K))).astype(np.complex64)
    return H, X


def _partial(H, X, cur, lo, hi):
    conv = np.zeros((N_OUT, K), np.complex64)
    for o in range(N_OUT):
        for i in range(N_IN):
            for seg in range(lo, hi):
                conv[o] += H[o, i, seg] * X[i, (cur + seg) % S]
    return conv


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H, X = _spectra(5)
    lo, hi = mimo_segment_range(S, rank, world)
    part = _partial(H, X, 4, lo, hi)
    t = torch.from_numpy(part.view(np.float32).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    np.save(os.path.join(outdir, f"conv{rank}.npy"), t.numpy().view(np.complex64))
    dist.destroy_process_group()


def test_segment_ranges_partition():
    for s in (0, 1, 7, 188, 938):
        for g in (1, 2, 3, 8):
            r = [mimo_segment_range(s, i, g) for i in range(g)]
            assert r[0][0] == 0 and r[-1][1] == s
            assert all(r[i][1] == r[i + 1][0] for i in range(g - 1))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1


def test_two_rank_allreduce_of_partial_spectra(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    H, X = _spectra(5)
    full = _partial(H, X, 4, 0, S)
    for r in range(world):
        got = np.load(tmp_path / f"conv{r}.npy").reshape(N_OUT, K)
        assert np.max(np.abs(got - full)) <= 1e-5 * np.max(np.abs(full))
