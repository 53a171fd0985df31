"""N > 1 host logic on CPU: two gloo ranks each own a disjoint channel range (no data-path
collective), process it independently, and only the timing goes through a max-reduction.
The per-rank arithmetic here is the CPU oracle standing in for the engine (no GPU in this suite)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bench
import oracle

CH, B, L, BLOCKS = 3, 64, 300, 6


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_output(c0, c1):
    irs = bench.synth_irs(c0, c1 - c0, 0, L)
    x = bench.synth_noise(c0, c1 - c0, 0, B * BLOCKS)
    y = np.zeros_like(x)
    for i in range(c1 - c0):
        conv = oracle.FFTConvolver.init(irs[i], B, L)
        conv.process(x[i], y[i])
    return y


def _worker(rank, world, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c0, c1 = bench.rank_channel_range(rank, world, CH)
    y = _shard_output(c0, c1)
    np.save(os.path.join(outdir, f"y{rank}.npy"), y)
    # timings: rank r pretends to have taken (r+1) ms; everyone must see the max
    ms, e2e = bench.reduce_max([float(rank + 1), 10.0 * (rank + 1)])
    assert (ms, e2e) == (float(world), 10.0 * world)
    v = bench.aggregate_value(world, CH, BLOCKS, B, ms)
    assert abs(v - world * CH * BLOCKS * B / 48000 / (world / 1000.0)) < 1e-9
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_channels_without_collective(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"y{r}.npy") for r in range(world)])
    ref = _shard_output(0, world * CH)  # one process owning every channel
    assert np.array_equal(got, ref)


def test_channel_ranges_partition_the_job():
    world, per = 8, 4096
    ranges = [bench.rank_channel_range(r, world, per) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == world * per
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        bench.rank_channel_range(8, 8, per)
