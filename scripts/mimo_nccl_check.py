#!/usr/bin/env python
"""Run under torchrun (>= 2 GPUs): IR-partition-sharded convolution matrix, partial spectra exchanged
by the NCCL all-reduce and by the NVLink peer exchange (reduce-kernel peer stores + flags), checked on
every rank against the unsharded engine and (rank 0) the CPU oracle.  Exit code 0 = parity."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bench
    import fft_convolution_b200 as F
    from fft_convolution_b200.distributed import ShardedMimoConvolver
    n_out, n_in, B, L, blocks = 4, 3, 128, 128 * 23 + 7, 30
    h = bench.synth_irs(0, n_out * n_in, 0, L).reshape(n_out, n_in, L)
    x = bench.synth_noise(50, n_in, 0, B * blocks)
    worst = 0.0
    for exchange in ("nccl", "peer"):
        sharded = ShardedMimoConvolver(h, B, L, device=local, exchange=exchange)
        whole = F.MimoConvolver.init(h, B, L, device=local)
        ref = None
        if rank == 0:
            from mimo_oracle import MimoOracle
            ref = MimoOracle(h, B, L)
        d_out = torch.empty((n_out, B), dtype=torch.float32, device=f"cuda:{local}")
        out_w = np.zeros((n_out, B), np.float32)
        for b in range(blocks):
            blk = np.ascontiguousarray(x[:, b * B:(b + 1) * B])
            d_in = torch.from_numpy(blk).cuda(local)
            sharded.process_dev(d_in, d_out)
            sharded.m.sync()
            got = d_out.cpu().numpy()
            whole.process(blk, out_w)
            r = float(np.sqrt(np.mean(out_w.astype(np.float64) ** 2)))
            worst = max(worst, float(np.max(np.abs(got - out_w))) / max(r, 0.05))
            if ref is not None:
                yo = ref.process(blk)
                worst = max(worst, float(np.max(np.abs(got - yo))) / max(r, 0.05))
        dist.barrier()
    t = torch.tensor([worst], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"mimo nccl check: world {world}, segments {sharded.m.seg_count}, worst err {float(t[0]):.3e} x RMS")
    dist.destroy_process_group()
    sys.exit(0 if float(t[0]) <= 2e-5 else 1)


if __name__ == "__main__":
    main()
