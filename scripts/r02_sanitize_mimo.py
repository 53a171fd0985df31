#!/usr/bin/env python
"""Small runs of the convolution-matrix paths for compute-sanitizer (memcheck / racecheck / synccheck): k_mac_rt in every CTA
shape, the tile kernel, the tensor-core K4, IR-partition shards with the peer exchange (all-gather and reduce-scatter form)
in one process.  No oracle here: this script only has to drive the kernels; a finite, non-zero output is the sanity check.
   compute-sanitizer --tool memcheck python scripts/r02_sanitize_mimo.py [rt|tile|tc|peer ...]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import fft_convolution_b200 as F  # noqa: E402
from fft_convolution_b200 import _lib  # noqa: E402

lib = _lib.load()


def irs(n_out, n_in, L):
    return bench.synth_irs(0, n_out * n_in, 0, L).reshape(n_out, n_in, L)


def run(g, n_rows_in, n_rows_out, B, nblocks):
    out = np.zeros((n_rows_out, B), np.float32)
    for b in range(nblocks):
        g.process(bench.synth_noise(0, n_rows_in, b * B, B), out)
    assert np.isfinite(out).all() and np.abs(out).max() > 0
    return float(np.abs(out).max())


def rt():
    for n_out, n_in, B, L, NS, nblocks in [(16, 3, 64, 64 * 25 + 1, 1, 29), (3, 2, 64, 64 * 13 + 5, 2, 16), (8, 3, 32, 32 * 9, 5, 12),
                                           (5, 1, 128, 128 * 6 + 1, 11, 8), (16, 2, 64, 64 * 10, 8, 13), (16, 3, 32, 32 * 37 + 5, 16, 41),
                                           (20, 2, 32, 32 * 5, 19, 8)]:
        g = F.MimoConvolver.init(irs(n_out, n_in, L), B, L, n_streams=NS, tensor_cores=False)
        assert g.mac_kernel == "register_tile"
        print("rt", (n_out, n_in, B, L, NS), run(g, NS * n_in, NS * n_out, B, nblocks), flush=True)
        g.close()


def tile():
    _lib.check(lib.fcb_tune(b"mimo_rt", 0))
    try:
        for n_out, n_in, B, L, NS, nblocks in [(4, 4, 128, 1500, 1, 14), (16, 2, 64, 64 * 10, 5, 12)]:
            g = F.MimoConvolver.init(irs(n_out, n_in, L), B, L, n_streams=NS, tensor_cores=False)
            assert g.mac_kernel == "tile"
            print("tile", (n_out, n_in, B, L, NS), run(g, NS * n_in, NS * n_out, B, nblocks), flush=True)
            g.close()
    finally:
        _lib.check(lib.fcb_tune(b"mimo_rt", 1))


def tc():
    for n_out, n_in, B, L, NS, nblocks in [(16, 2, 32, 32 * 37 + 5, 3, 40), (16, 3, 32, 32 * 20, 40, 22)]:
        g = F.MimoConvolver.init(irs(n_out, n_in, L), B, L, n_streams=NS, tensor_cores=True)
        assert g.mac_kernel == "tensor"
        print("tc", (n_out, n_in, B, L, NS), run(g, NS * n_in, NS * n_out, B, nblocks), flush=True)
        g.close()


def peer():
    """all shards of a job in one process on one GPU: every publish before any K3 (tests/test_gpu_mimo.py)"""
    import torch
    n_out, n_in, B, L, NS, shards, nblocks = 4, 2, 64, 64 * 13 + 5, 2, 2, 16
    for scatter in (False, True):
        parts = [F.MimoConvolver.init(irs(n_out, n_in, L), B, L, n_streams=NS, shard_index=g, shard_count=shards, tensor_cores=False)
                 for g in range(shards)]
        inboxes = [p.peer_inbox() for p in parts]
        for p in parts:
            p.peer_attach_ptrs(inboxes)
            if scatter:
                p.peer_set_scatter(True)
        d_in = torch.empty((NS * n_in, B), dtype=torch.float32, device="cuda")
        d_out = [torch.zeros((NS * n_out, B), dtype=torch.float32, device="cuda") for _ in parts]
        for b in range(nblocks):
            d_in.copy_(torch.from_numpy(bench.synth_noise(0, NS * n_in, b * B, B)))
            torch.cuda.synchronize()
            for p in parts:
                p.partial_dev(d_in.data_ptr(), B)
            for p in parts:
                p.sync()
            for p, o in zip(parts, d_out):
                p.finish_dev(o.data_ptr(), B)
            for p in parts:
                p.sync()
        print("peer", "reduce-scatter" if scatter else "all-gather", float(d_out[0].abs().max()), flush=True)
        for p in parts:
            p.close()


if __name__ == "__main__":
    for name in (sys.argv[1:] or ["rt", "tile", "tc"]):
        {"rt": rt, "tile": tile, "tc": tc, "peer": peer}[name]()
    print("sanitize mimo: done")
