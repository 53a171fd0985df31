#!/usr/bin/env python
"""Absolute numbers for BASELINE configs[4]: 16 x 16 convolution matrix, 10 s (480 000-tap) IRs,
block 512.  One process per GPU; with WORLD_SIZE > 1 the IR is sharded by partition and the
partial spectra go through one NCCL all-reduce per block.  Prints one JSON line (rank 0).
Not a bench.py line (the headline metric is the independent-channel config); kept under scripts/."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402  (synthetic generators)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16)
    ap.add_argument("--block", type=int, default=512)
    ap.add_argument("--ir-seconds", type=float, default=10.0)
    ap.add_argument("--streams", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="how IR-partition shards exchange partial spectra")
    ap.add_argument("--tc", type=int, default=-1, help="1/0 force the tensor-core matrix MAC (K4) on/off, -1 library default")
    ap.add_argument("--rt", type=int, default=1, help="0: the shared-memory tile kernel instead of the register-tiled MAC (k_mac_rt) below the tensor-core threshold")
    ap.add_argument("--tune", default="", help="extra fcb_tune settings, e.g. mimo_rt_wb=1,mimo_tc_min=40")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local))
    from fft_convolution_b200.distributed import ShardedMimoConvolver
    from fft_convolution_b200 import _lib
    lib = _lib.load()
    _lib.check(lib.fcb_tune(b"mimo_rt", a.rt))
    for kv in filter(None, a.tune.split(",")):
        k, v = kv.split("=")
        _lib.check(lib.fcb_tune(k.encode(), int(v)))
    N, B, L, NS = a.n, a.block, int(a.ir_seconds * 48000), a.streams
    h = bench.synth_irs(0, N * N, 0, L).reshape(N, N, L)  # IR index c = out*N + in
    m = ShardedMimoConvolver(h, B, L, n_streams=NS, device=local, tensor_cores=None if a.tc < 0 else bool(a.tc), exchange=a.exchange)
    S = m.m.seg_count
    lo, hi = m.m.segment_range
    x = [torch.from_numpy(bench.synth_noise(0, NS * N, B * i, B)).cuda(local) for i in range(8)]
    out = torch.empty((NS * N, B), dtype=torch.float32, device=f"cuda:{local}")
    for i in range(a.warmup):
        m.process_dev(x[i % 8], out)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    _lib.check(lib.fcb_profile_mac(1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(m.stream)
    for i in range(a.steps):
        m.process_dev(x[i % 8], out)
    e1.record(m.stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    import ctypes as C
    tot, nl = C.c_double(), C.c_uint64()
    _lib.check(lib.fcb_profile_mac_read(C.byref(tot), C.byref(nl)))
    _lib.check(lib.fcb_profile_mac(0))
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    k2_ms = max(tot.value / max(nl.value, 1), 1e-9)
    K = B + 1
    rows = hi - max(lo, 1)
    ir_bytes = N * N * rows * K * 8  # IR matrix rows this shard streams from HBM per block
    ring_bytes = NS * N * N * rows * K * 8  # ring rows, re-read per output (L2 resident: 16 rings = 61 MB)
    if rank == 0:
        print(json.dumps({
            "config": f"MIMO {N}x{N}, IR {a.ir_seconds:g} s ({L} taps, S={S}), block {B}, streams {NS}, {world} GPU(s) (IR-partition shards)",
            "tensor_cores": bool(m.m.uses_tensor_cores), "mac_kernel": m.m.mac_kernel, "exchange": m.exchange,
            "T_cmac_per_s": NS * N * N * (hi - lo) * B / (float(t[0]) / 1e3) / 1e12,
            "ms_per_block": float(t[0]), "block_period_ms": 1000.0 * B / 48000, "realtime_factor": 1000.0 * B / 48000 / float(t[0]),
            "k2_ms": k2_ms, "k2_ir_GBs": ir_bytes / (k2_ms / 1e3) / 1e9, "k2_ir_plus_ring_GBs": (ir_bytes + ring_bytes) / (k2_ms / 1e3) / 1e9,
            "ir_bytes_per_block_this_shard": ir_bytes, "segments_this_shard": [lo, hi],
            "allreduce_bytes": 2 * 4 * NS * N * B if world > 1 else 0,
        }), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
