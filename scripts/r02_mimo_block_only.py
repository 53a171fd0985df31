#!/usr/bin/env python
"""bench.py's `mimo` block alone (BASELINE configs[4] sharded by IR partition over the N GPUs; 1 / 16 / 128 streams; peer and
NCCL exchange, all-gather and reduce-scatter form; every entry checked against the unsharded engine):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/r02_mimo_block_only.py
Rank 0 prints one JSON line."""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
saved = os.dup(1)
os.dup2(2, 1)  # the NCCL banner goes to stderr
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dist.barrier()
torch.cuda.synchronize()
sys.stdout.flush()
os.dup2(saved, 1)
res = bench.run_mimo(local, rank, world, int(os.environ.get("STEPS", 200)), 20)
if rank == 0:
    print(json.dumps({"n_gpus": world, "mimo": res}), flush=True)
dist.barrier()
dist.destroy_process_group()
