"""Convolution matrix (BASELINE configs[4] shape, reduced sizes) against OUT x IN oracle
convolvers; IR-partition shards emulated on one GPU: partial spectra summed on the host exactly
as the NCCL all-reduce would (one kernel sequence per shard, never concurrent waiting kernels)."""
import numpy as np
import pytest

import oracle
from mimo_oracle import MimoOracle
from refsignals import rms, WholeRun

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def F():
    import fft_convolution_b200 as f
    return f


def _irs(n_out, n_in, L, upd=0):
    return np.stack([np.stack([oracle.gen_ir(o * n_in + i, upd, L) for i in range(n_in)]) for o in range(n_out)])


@pytest.mark.parametrize("n_out,n_in,B,L", [(3, 2, 64, 700), (4, 4, 128, 1500), (16, 16, 32, 200), (1, 1, 256, 1000)])
def test_mimo_matches_oracle(F, n_out, n_in, B, L):
    h = _irs(n_out, n_in, L)
    nblocks = 14
    x = np.stack([oracle.gen_noise(100 + i, 0, B * nblocks) for i in range(n_in)])
    g, o = F.MimoConvolver.init(h, B, L), MimoOracle(h, B, L)
    out = np.zeros((n_out, B), np.float32)
    run = WholeRun()
    for b in range(nblocks):
        blk = np.ascontiguousarray(x[:, b * B:(b + 1) * B])
        g.process(blk, out)
        run.add(out, o.process(blk))
    run.check(TOL, "matrix vs OUT x IN oracle convolvers")


def test_mimo_streams_share_the_matrix(F):
    n_out, n_in, B, L, NS = 2, 3, 64, 500, 3
    h = _irs(n_out, n_in, L)
    x = np.stack([oracle.gen_noise(200 + i, 0, B * 10) for i in range(NS * n_in)])
    g = F.MimoConvolver.init(h, B, L, n_streams=NS)
    singles = [F.MimoConvolver.init(h, B, L) for _ in range(NS)]
    out = np.zeros((NS * n_out, B), np.float32)
    one = np.zeros((n_out, B), np.float32)
    run = WholeRun()
    for b in range(10):
        blk = np.ascontiguousarray(x[:, b * B:(b + 1) * B])
        g.process(blk, out)
        for s in range(NS):
            singles[s].process(blk[s * n_in:(s + 1) * n_in], one)
            run.add(out[s * n_out:(s + 1) * n_out], one)
    # several streams run the register-tiled MAC (k_mac_rt), one stream the tile kernel: same sums, different association
    run.check(TOL, "streams sharing the matrix vs one engine per stream")


@pytest.mark.parametrize("shards", [2, 3, 8])
def test_ir_partition_shards_sum_to_the_whole(F, shards):
    """every shard owns S*g/G..S*(g+1)/G of the IR segments; partial spectra summed (as the NCCL
    all-reduce does) then K3 on every shard == the unsharded engine up to f32 summation order"""
    import torch
    from fft_convolution_b200.distributed import _DeviceBuffer
    n_out, n_in, B, L = 3, 2, 64, 64 * 11 + 5
    h = _irs(n_out, n_in, L)
    x = np.stack([oracle.gen_noise(300 + i, 0, B * 16) for i in range(n_in)])
    whole = F.MimoConvolver.init(h, B, L)
    parts = [F.MimoConvolver.init(h, B, L, shard_index=g, shard_count=shards) for g in range(shards)]
    S = whole.seg_count
    ranges = [p.segment_range for p in parts]
    assert ranges == [F.mimo_segment_range(S, g, shards) for g in range(shards)]
    assert ranges[0][0] == 0 and ranges[-1][1] == S and all(ranges[i][1] == ranges[i + 1][0] for i in range(shards - 1))
    ref_o = MimoOracle(h, B, L)
    out_w = np.zeros((n_out, B), np.float32)
    d_in = torch.empty((n_in, B), dtype=torch.float32, device="cuda")
    d_out = torch.empty((n_out, B), dtype=torch.float32, device="cuda")
    run_ref, run_whole = WholeRun(), WholeRun()
    for b in range(16):
        blk = np.ascontiguousarray(x[:, b * B:(b + 1) * B])
        whole.process(blk, out_w)
        d_in.copy_(torch.from_numpy(blk))
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
        outs = []
        for p in parts:
            p.finish_dev(d_out.data_ptr(), B)
            p.sync()
            outs.append(d_out.cpu().numpy().copy())
        for o in outs[1:]:
            assert np.array_equal(o, outs[0])  # every shard ends with the same block
        run_ref.add(outs[0], ref_o.process(blk))
        run_whole.add(outs[0], out_w)
    run_ref.check(TOL, "sharded matrix vs oracle")
    run_whole.check(TOL, "sharded vs unsharded matrix")


@pytest.mark.parametrize("shards,tc", [(2, False), (3, False), (4, True)])
def test_peer_exchange_shards_in_one_process(F, shards, tc):
    """the NVLink peer exchange (reduce-kernel epilogue stores into every shard's inbox + release flag,
    K3 acquires G flags and sums G slots) with all shards of a job on one GPU: every publish is issued
    before any K3, so no kernel ever waits on a later one.  All shards end bit-identical and == oracle."""
    import torch
    n_out, n_in, B, L, NS = (16, 2, 64, 64 * 19 + 5, 3) if tc else (3, 2, 64, 64 * 11 + 5, 2)
    h = _irs(n_out, n_in, L)
    nblocks = 26
    x = np.stack([oracle.gen_noise(900 + i, 0, B * nblocks) for i in range(NS * n_in)])
    parts = [F.MimoConvolver.init(h, B, L, n_streams=NS, shard_index=g, shard_count=shards, tensor_cores=tc)
             for g in range(shards)]
    assert all(p.uses_tensor_cores == tc for p in parts)
    inboxes = [p.peer_inbox() for p in parts]
    for p in parts:
        p.peer_attach_ptrs(inboxes)
    refs = [MimoOracle(h, B, L) for _ in range(NS)]
    d_in = torch.empty((NS * n_in, B), dtype=torch.float32, device="cuda")
    d_out = [torch.empty((NS * n_out, B), dtype=torch.float32, device="cuda") for _ in parts]
    run = WholeRun()
    for b in range(nblocks):
        blk = np.ascontiguousarray(x[:, b * B:(b + 1) * B])
        d_in.copy_(torch.from_numpy(blk))
        torch.cuda.synchronize()
        for p in parts:
            p.partial_dev(d_in.data_ptr(), B)
        for p in parts:
            p.sync()
        for p, o in zip(parts, d_out):
            p.finish_dev(o.data_ptr(), B)
        for p in parts:
            p.sync()
        outs = [o.cpu().numpy() for o in d_out]
        for o in outs[1:]:
            assert np.array_equal(o, outs[0])
        for s_ in range(NS):
            run.add(outs[0][s_ * n_out:(s_ + 1) * n_out], refs[s_].process(blk[s_ * n_in:(s_ + 1) * n_in]))
    run.check(TOL, "peer-exchange shards vs oracle")


@pytest.mark.parametrize("shards,tc,NS,n_out", [(2, False, 2, 3), (3, False, 1, 4), (4, True, 3, 16), (3, True, 5, 16)])
def test_peer_exchange_reduce_scatter_form(F, shards, tc, NS, n_out):
    """reduce-scatter form of the peer exchange: a partial row travels to its owner only and every shard finishes just
    its own rows [R g / G, R (g+1) / G) of the R = NS * OUT output rows (uneven splits included); the rows of all shards
    together are the full result.  All shards in one process on one GPU, every publish issued before any K3."""
    import torch
    n_in, B = 2, 64
    L = B * (19 if tc else 11) + 5
    h = _irs(n_out, n_in, L)
    nblocks = 24
    x = np.stack([oracle.gen_noise(700 + i, 0, B * nblocks) for i in range(NS * n_in)])
    parts = [F.MimoConvolver.init(h, B, L, n_streams=NS, shard_index=g, shard_count=shards, tensor_cores=tc) for g in range(shards)]
    inboxes = [p.peer_inbox() for p in parts]
    for p in parts:
        p.peer_attach_ptrs(inboxes)
        p.peer_set_scatter(True)
    rows = [p.owned_rows for p in parts]
    R = NS * n_out
    assert rows[0][0] == 0 and rows[-1][1] == R and all(rows[i][1] == rows[i + 1][0] for i in range(shards - 1))
    refs = [MimoOracle(h, B, L) for _ in range(NS)]
    d_in = torch.empty((NS * n_in, B), dtype=torch.float32, device="cuda")
    d_out = torch.full((R, B), float("nan"), dtype=torch.float32, device="cuda")  # ONE buffer: every shard writes its rows
    run = WholeRun()
    for b in range(nblocks):
        blk = np.ascontiguousarray(x[:, b * B:(b + 1) * B])
        d_in.copy_(torch.from_numpy(blk))
        d_out.fill_(float("nan"))
        torch.cuda.synchronize()
        for p in parts:
            p.partial_dev(d_in.data_ptr(), B)
        for p in parts:
            p.sync()
        for p in parts:
            p.finish_dev(d_out.data_ptr(), B)
        for p in parts:
            p.sync()
        out = d_out.cpu().numpy()
        assert not np.isnan(out).any()  # the shards' rows cover the output exactly
        for s_ in range(NS):
            run.add(out[s_ * n_out:(s_ + 1) * n_out], refs[s_].process(blk[s_ * n_in:(s_ + 1) * n_in]))
    run.check(TOL, "reduce-scatter peer exchange vs oracle")


@pytest.mark.parametrize("shards,tc,scatter,NS,n_out", [(2, False, False, 2, 3), (3, False, True, 4, 6), (2, True, False, 3, 16)])
def test_peer_exchange_overlapped_finish(F, shards, tc, scatter, NS, n_out):
    """fcb_mimo_set_overlap: K3 on its own stream, ordered by events only (K3 after its block's reduce, the next block's
    reduce after K3).  All shards in one process on one GPU — partials, a sync, finishes, a sync per block as that setting
    demands — so this pins the event plumbing and the join / sync / reset semantics; blocks in flight across GPUs are
    checked by bench.py's `mimo` block (every `_overlap` entry against the unsharded engine)."""
    import torch
    n_in, B = 2, 64
    L = B * (19 if tc else 11) + 5
    h = _irs(n_out, n_in, L)
    nblocks = 24
    x = np.stack([oracle.gen_noise(800 + i, 0, B * nblocks) for i in range(NS * n_in)])
    parts = [F.MimoConvolver.init(h, B, L, n_streams=NS, shard_index=g, shard_count=shards, tensor_cores=tc) for g in range(shards)]
    inboxes = [p.peer_inbox() for p in parts]
    for p in parts:
        p.peer_attach_ptrs(inboxes)
        if scatter:
            p.peer_set_scatter(True)
        p.set_overlap(True)
    refs = [MimoOracle(h, B, L) for _ in range(NS)]
    R = NS * n_out
    d_in = torch.empty((NS * n_in, B), dtype=torch.float32, device="cuda")
    d_out = [torch.empty((R, B), dtype=torch.float32, device="cuda") for _ in parts]
    if scatter:
        d_out = [d_out[0]] * shards  # ONE buffer: every shard writes its own rows
    run = WholeRun()
    for b in range(nblocks):
        if b == 15:  # reset in the middle: joins the finish stream before it clears the overlap
            for p in parts:
                p.reset()
            refs = [MimoOracle(h, B, L) for _ in range(NS)]
        blk = np.ascontiguousarray(x[:, b * B:(b + 1) * B])
        d_in.copy_(torch.from_numpy(blk))
        torch.cuda.synchronize()
        for p in parts:
            p.partial_dev(d_in.data_ptr(), B)
        for p in parts:
            p.sync()
        for p, o in zip(parts, d_out):
            p.finish_dev(o.data_ptr(), B)
        for p in parts:
            p.join()
        for p in parts:
            p.sync()
        outs = [o.cpu().numpy() for o in d_out]
        if not scatter:
            for o in outs[1:]:
                assert np.array_equal(o, outs[0])
        for s_ in range(NS):
            run.add(outs[0][s_ * n_out:(s_ + 1) * n_out], refs[s_].process(blk[s_ * n_in:(s_ + 1) * n_in]))
    run.check(TOL, "peer exchange with overlapped finish vs oracle")



def test_nccl_sharded_mimo_two_gpus():
    """IR-partition shards on 2 GPUs, exchanged by NCCL all-reduce and by the NVLink peer exchange (skipped on a 1-GPU box)"""
    import subprocess
    import sys
    from pathlib import Path
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(root / "scripts" / "mimo_nccl_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "mimo nccl check" in r.stdout


@pytest.mark.parametrize("n_streams,B", [(1, 512), (5, 64), (4, 1024)])
def test_tile_kernel_matches_generic_k2(F, n_streams, B):
    """matrix K2 with in-CTA reuse (k_mac_tile) vs the per-channel K2: same sums, different
    association — equal to f32 rounding, and both within tolerance of the oracle"""
    from fft_convolution_b200 import _lib
    lib = _lib.load()
    n_out, n_in, L = 5, 3, B * 21 + 3
    h = _irs(n_out, n_in, L)
    x = np.stack([oracle.gen_noise(400 + i, 0, B * 8) for i in range(n_streams * n_in)])
    outs = {}
    _lib.check(lib.fcb_tune(b"mimo_rt", 0))  # several streams would otherwise run k_mac_rt (tests/test_gpu_mimo_rt.py)
    for tile in (1, 0):
        _lib.check(lib.fcb_tune(b"mimo_tile", tile))
        g = F.MimoConvolver.init(h, B, L, n_streams=n_streams)
        assert g.mac_kernel == "tile"
        y = np.zeros((n_streams * n_out, B * 8), np.float32)
        blk_out = np.zeros((n_streams * n_out, B), np.float32)
        for b in range(8):
            g.process(np.ascontiguousarray(x[:, b * B:(b + 1) * B]), blk_out)
            y[:, b * B:(b + 1) * B] = blk_out
        outs[tile] = y
    _lib.check(lib.fcb_tune(b"mimo_tile", 1))
    _lib.check(lib.fcb_tune(b"mimo_rt", 1))
    ref = MimoOracle(h, B, L).process(x[:n_in])
    r = rms(ref)
    assert np.max(np.abs(outs[1][:n_out] - ref)) <= TOL * r
    assert np.max(np.abs(outs[1] - outs[0])) <= TOL * r
