"""Multi-GPU plumbing (torch.distributed is plumbing only; all arithmetic is in the CUDA library).

* Independent channels: every rank owns a disjoint channel range and its own FFTConvolver batch —
  no collective on the data path (bench.py does exactly this).
* Convolution matrix with very long IRs: the IR is sharded by partition (contiguous ranges of IR
  segments); each rank computes the partial spectra of its range.  exchange="peer" (default on one
  node): the reduce kernel's epilogue stores them into every rank's inbox over NVLink and K3 sums the
  G inboxes after an acquire on G flags — no collective kernel.  exchange="nccl": one NCCL all-reduce
  per block (16 x B complex = 64 KB at B = 512), after which every rank runs K3.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .convolvers import MimoConvolver


class _DeviceBuffer:
    """exposes a raw device pointer of the C library to torch through __cuda_array_interface__"""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class ShardedMimoConvolver:
    """One IR-partition shard per rank of the default process group (NCCL on GPUs)."""

    def __init__(self, responses, block_size: int, max_response_length: int, *, n_streams: int = 1, device: int = 0,
                 tensor_cores: bool | None = None, exchange: str = "peer", scatter: bool = False, overlap: bool = False):
        """overlap=True (peer exchange): K3 of a block — the kernel that waits for the peers — runs beside K1 and the MAC
        of the next block (throughput of back-to-back blocks); `out` is ordered on `self.stream` by join(), for the host by
        sync().
        scatter=True: reduce-scatter instead of all-gather / all-reduce — every rank FINISHES only its own output rows
        (`self.rows`), the result is sharded by row over the ranks (1/G of the exchange bytes, 1/G of the K3 work)"""
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.m = MimoConvolver.init(responses, block_size, max_response_length, n_streams=n_streams,
                                    shard_index=self.rank, shard_count=self.world, device=device,
                                    stream=self.stream.cuda_stream, tensor_cores=tensor_cores)
        ptr, n = self.m.conv_buffer()
        self._keep = _DeviceBuffer(ptr, n)
        self.conv = torch.as_tensor(self._keep, device=f"cuda:{device}")
        self.B = self.m.block_size
        self.exchange = exchange if self.world > 1 else "none"
        self.scatter = bool(scatter) and self.world > 1
        self.rows = self.m.owned_rows if self.scatter else (0, n_streams * self.m.n_out)
        if self.scatter and self.exchange == "nccl" and (n_streams * self.m.n_out) % self.world:
            raise ValueError("reduce_scatter needs the output rows to divide evenly over the ranks")
        self.overlap = bool(overlap) and self.exchange == "peer"
        if self.exchange == "peer":
            if self.scatter:
                self.m.peer_set_scatter(True)
            if self.overlap:
                self.m.set_overlap(True)
            mine = torch.frombuffer(bytearray(self.m.peer_export()), dtype=torch.uint8).to(f"cuda:{device}")
            every = torch.empty((self.world, 64), dtype=torch.uint8, device=f"cuda:{device}")
            dist.all_gather_into_tensor(every, mine)
            self.m.peer_attach([bytes(row.tolist()) for row in every.cpu()])
            dist.barrier()

    def process_dev(self, x: torch.Tensor, out: torch.Tensor) -> None:
        """x: [NS*IN, B] float32 on this rank's GPU (every rank is given the same block);
        out: [NS*OUT, B].  Every rank ends with the full result."""
        with torch.cuda.stream(self.stream):
            self.m.partial_dev(x.data_ptr(), x.stride(0))
            if self.exchange == "nccl" and self.scatter:
                lo, hi = self.rows
                mine = self.conv[lo * 2 * self.B:hi * 2 * self.B]  # my rows of the conv buffer, reduced in place
                dist.reduce_scatter_tensor(mine, self.conv, op=dist.ReduceOp.SUM)
                self.m.finish_rows_dev(out.data_ptr(), out.stride(0), lo, hi)
                return
            if self.exchange == "nccl":
                dist.all_reduce(self.conv, op=dist.ReduceOp.SUM)
            self.m.finish_dev(out.data_ptr(), out.stride(0))

    def join(self) -> None:
        """overlap=True: everything queued on self.stream after this call sees the outputs of every block so far"""
        self.m.join()

    def sync(self) -> None:
        """host-side wait for every block so far; raises if the peer exchange timed out"""
        self.m.sync()
