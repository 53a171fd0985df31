#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json metric).

Workload (config.workload): BASELINE configs[3] — independent 48 kHz channels, 2 s (96 000-tap)
impulse response per channel (each its own IR), block 512, one FFTConvolver per channel,
4096 channels per GPU (weak scaling: every rank owns its own 4096 channels, no collective).
One "step" = one 512-sample block for every channel (K1 -> K2 -> K3).

  python bench.py [--gpus N --steps K --warmup W]            our arm (CUDA, sm_100a)
  python bench.py --impl reference [...]                     CPU restatement of the reference
                                                             algorithm on all host cores

Prints ONE JSON line (rank 0).  Synthetic data per SURVEY.md §8(d) (splitmix64 white noise,
random-decay unit-energy IRs), generated here with numpy — the product arm never touches oracle/.

Beside the headline (`value`, `e2e`, `roofline`, `cpu_baseline`, unchanged in meaning) the line carries:
  sustained   the same step timed over a >= 1 s window (the driver's K may be 20 steps = 18 ms)
  strong      BASELINE configs[3] read as "4096 channels sharded across N GPUs": 4096 / N channels per GPU
  realtime    the >= 1e5-channel claim measured, not extrapolated: 12 500 channels per GPU (1e5 / 8) and the largest
              channel count whose block still fits the 10.667 ms period, 1000 consecutive blocks each through the
              end-to-end host path, with p50 / p99 / max block time and the deadlines missed
  mimo        (N > 1) BASELINE configs[4]: the 16 x 16 matrix with 10 s IRs sharded by IR partition over the N GPUs,
              1 and 128 streams, peer exchange and NCCL all-reduce, checked in-run against the unsharded engine
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

SAMPLE_RATE = 48000
METRIC = "real-time 48 kHz channel count (channel-sec/sec) at IR 2 s, block 512"
UNIT = "channel-sec/sec"


# ---------------------------------------------------------------------------------------------
# synthetic data — numpy, bit-identical to the generators the tests use
# ---------------------------------------------------------------------------------------------
def mix64(v: np.ndarray) -> np.ndarray:
    z = v + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def synth_noise(chan0: int, nchan: int, first_sample: int, n: int) -> np.ndarray:
    c = (np.arange(chan0, chan0 + nchan, dtype=np.uint64) << np.uint64(32))[:, None]
    i = np.arange(first_sample, first_sample + n, dtype=np.uint64)[None, :]
    r = mix64(np.uint64(0x5EED0001) + c + i)
    u = (r >> np.uint64(40)).astype(np.float32) / np.float32(16777216.0)
    return (np.float32(2.0) * u - np.float32(1.0)).astype(np.float32)


def synth_irs(chan0: int, nchan: int, update_index: int, length: int, chunk: int = 64) -> np.ndarray:
    out = np.empty((nchan, length), np.float32)
    decay = np.exp(-6.9078 * np.arange(length, dtype=np.float64) / float(length))[None, :]
    i = np.arange(length, dtype=np.uint64)[None, :]
    seed = np.uint64(0x5EED0002) + (np.uint64(update_index) << np.uint64(48))
    for c0 in range(0, nchan, chunk):
        c1 = min(nchan, c0 + chunk)
        c = (np.arange(chan0 + c0, chan0 + c1, dtype=np.uint64) << np.uint64(32))[:, None]
        r = mix64(seed + c + i)
        u = (r >> np.uint64(40)).astype(np.float64) / 16777216.0
        v = (2.0 * u - 1.0) * decay
        e = np.sum(v * v, axis=1, keepdims=True)
        out[c0:c1] = (v / np.sqrt(e)).astype(np.float32)
    return out


# ---------------------------------------------------------------------------------------------
# multi-GPU plan: channels are independent, so ranks own disjoint contiguous channel ranges and
# the data path has no collective; the only exchange is the max over ranks of the timings.
def rank_channel_range(rank: int, world: int, channels_per_gpu: int) -> tuple[int, int]:
    """global [first, last) channel indices owned by `rank` (weak scaling: fixed per-GPU count)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    return rank * channels_per_gpu, (rank + 1) * channels_per_gpu


def reduce_max(values, device=None):
    """max over ranks of a list of floats (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def aggregate_value(world: int, channels_per_gpu: int, steps: int, block: int, ms_max: float) -> float:
    """whole-job channel-seconds of audio per wall second"""
    return world * channels_per_gpu * steps * block / SAMPLE_RATE / (ms_max / 1000.0)


# ---------------------------------------------------------------------------------------------
def measured_peak_gbs() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent clocks via NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_ev = index, [], set(), None, threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_ev.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_ev.wait(0.01)

    def stop(self) -> dict:
        self._stop_ev.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
_CPU_DATA: dict = {}


def host_threads() -> int:
    """all host cores this process may use (torchrun exports OMP_NUM_THREADS=1, so ask the OS)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_port_run(channels: int, block: int, ir_len: int, calls: int, threads: int) -> tuple[float, float]:
    """CPU restatement of the reference algorithm (oracle/fftconv_oracle.c), one convolver per
    channel, channels over `threads` host threads; returns (channel-sec/sec, seconds of the block
    loop).  The synthetic data is generated once per shape and reused by later steps."""
    import oracle  # cpu_baseline / --impl reference legs only
    lib = oracle.load().lib
    key = (channels, block, ir_len, calls)
    if key not in _CPU_DATA:
        _CPU_DATA.clear()
        _CPU_DATA[key] = (synth_irs(1 << 20, channels, 0, ir_len), synth_noise(1 << 20, channels, 0, block * calls))
    irs, x = _CPU_DATA[key]
    out = np.empty_like(x)
    secs = lib.orc_batch_fftconv_run(channels, block, ir_len, irs, x, out, block, calls, threads)
    return channels * calls * block / SAMPLE_RATE / secs, secs


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    threads = host_threads()
    ir_len = int(args.ir_seconds * SAMPLE_RATE)
    channels = min(max(threads * 16, 16), 512)  # 16 channels per host thread, like the cpu_baseline leg
    calls = 94  # ~1 s of audio per channel per step
    # warm-up + timed steps, each step a bounded sample of the workload
    for _ in range(min(args.warmup, 3)):
        cpu_port_run(channels, args.block, ir_len, calls, threads)
    t_tot, work = 0.0, 0.0
    for _ in range(args.steps):
        v, secs = cpu_port_run(channels, args.block, ir_len, calls, threads)
        t_tot += secs
        work += channels * calls * args.block / SAMPLE_RATE
    value = work / t_tot
    sample = (f"{channels} of the workload's {args.channels} channels x {calls} blocks of {args.block} per step, {args.steps} steps; "
              "throughput is per channel, so the figure stands for the whole workload (working set "
              f"{channels * 2 * 188 * 513 * 8 / 1e6:.0f} MB >> LLC)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "CPU restatement of the reference algorithm (oracle/), not rustfft: no Rust toolchain in the image"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_dict(args):
    """the workload both arms are quoted on — the same dict on the CUDA arm and on the reference arm (what each arm
    actually timed per step is in `cpu_baseline.sample` / `tuning`, not here)"""
    return {"workload": f"FFTConvolver x {args.channels} independent channels per GPU, IR {args.ir_seconds:g} s "
                        f"({int(args.ir_seconds * SAMPLE_RATE)} taps, one IR per channel), block {args.block}, 48 kHz "
                        "(BASELINE configs[3])",
            "channels_per_gpu": args.channels, "block": args.block, "ir_taps": int(args.ir_seconds * SAMPLE_RATE),
            "l2": "working set 6.3 GB/GPU per step >> 126 MB L2, no flush needed"}



# ---------------------------------------------------------------------------------------------
def kernel_source_hash() -> str:
    """sha256 of the kernel sources the dominant kernel is built from: an ncu traffic figure is only quoted while the
    kernels it was captured from are the ones being timed (scripts/capture_traffic.sh re-captures and re-stamps it)"""
    import hashlib
    h = hashlib.sha256()
    for name in ("fused_kernel.cuh", "mac_kernels.cuh", "fft_kernels.cuh", "common.cuh"):
        h.update((ROOT / "fft_convolution_b200" / "csrc" / name).read_bytes())
    return h.hexdigest()[:16]


def bind_to_gpu_numa_node(local: int) -> dict:
    """Pin this rank's threads (and so the first-touch placement of the pinned buffers it allocates next) to the NUMA
    node its GPU hangs off, when the box exposes more than one node.  Returns what was found, for the bench line."""
    info = {"nodes_visible": None, "gpu_node": None, "bound": False}
    try:
        import torch
        nodes = sorted(int(p.name[4:]) for p in Path("/sys/devices/system/node").glob("node[0-9]*"))
        info["nodes_visible"] = len(nodes)
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int((Path("/sys/bus/pci/devices") / bdf / "numa_node").read_text())
        info["gpu_node"] = node
        if node >= 0 and len(nodes) > 1:
            cpus = set()
            for part in (Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(",")):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["bound"] = True
    except Exception as e:  # sysfs not there (containers): nothing to bind
        info["error"] = type(e).__name__
    return info


def percentile_ms(v, q):
    return float(np.percentile(np.asarray(v), q) * 1000.0)


def run_realtime(F, lib, local, rank, world, chan0, Cn_head, ms_e2e_per_block_head, B, L, blocks, barrier):
    """The real-time claim, measured: C channels per GPU through the END-TO-END host path (pinned host buffers in and
    out, synchronous call per block) for `blocks` consecutive blocks; every block's wall time against the block period.
    Two sizes: 12 500 channels (1e5 / 8 GPUs) and the largest count whose block is expected to fit the period.  The IRs
    of channel c are those of channel c mod 4096 of this rank (each channel still owns and streams its own copy)."""
    import torch
    from fft_convolution_b200 import _lib
    period = B / SAMPLE_RATE
    free_b, _ = torch.cuda.mem_get_info(local)
    per_ch = 2 * ((L + B - 1) // B) * B * 8 + 64 * B
    c_fit = int(0.85 * period * 1000.0 / ms_e2e_per_block_head * Cn_head) // 256 * 256  # 15 % headroom for host jitter
    c_mem = int(0.85 * free_b / per_ch) // 256 * 256
    sizes = [("1e5_over_8_gpus", 12500), ("largest_that_fits_the_period", max(256, min(c_fit, c_mem)))]
    base = synth_irs(chan0, min(4096, max(s for _, s in sizes)), 0, L)
    out = {}
    for name, Cn in sizes:
        if Cn * per_ch > 0.9 * free_b:
            out[name] = {"channels_per_gpu": Cn, "skipped": "does not fit the free HBM"}
            continue
        conv = F.FFTConvolver.init(np.zeros((Cn, 1), np.float32), B, L, device=local)
        eng = conv.engine
        for c0 in range(0, Cn, base.shape[0]):  # K5 in slabs of 4096 channels (layer 1: any channel range)
            n = min(base.shape[0], Cn - c0)
            _lib.check(lib.fcb_engine_set_ir(eng, c0, n, base.ctypes.data_as(C.c_void_p), L, L, 0))
        conv.sync()
        nbytes = Cn * B * 4
        p_in, p_out = lib.fcb_host_alloc(nbytes), lib.fcb_host_alloc(nbytes)
        h_in = np.ctypeslib.as_array(C.cast(p_in, C.POINTER(C.c_float)), shape=(Cn, B))
        h_out = np.ctypeslib.as_array(C.cast(p_out, C.POINTER(C.c_float)), shape=(Cn, B))
        h_in[:] = np.tile(synth_noise(chan0, min(Cn, 1024), 0, B), ((Cn + 1023) // 1024, 1))[:Cn]
        for _ in range(20):
            conv.process(h_in, h_out)
        barrier()
        times = np.empty(blocks)
        import gc
        gc.disable()  # an audio thread does not collect garbage; neither should the loop that stands in for one
        try:
            t_all = time.perf_counter()
            for i in range(blocks):
                t0 = time.perf_counter()
                conv.process(h_in, h_out)
                times[i] = time.perf_counter() - t0
            t_all = time.perf_counter() - t_all
        finally:
            gc.enable()
        barrier()
        stats = reduce_max([percentile_ms(times, 50), percentile_ms(times, 99), float(times.max() * 1000.0),
                            float((times > period).sum()), t_all * 1000.0], device=f"cuda:{local}")
        out[name] = {"channels_per_gpu": Cn, "channels_all_gpus": Cn * world, "blocks": blocks,
                     "block_period_ms": period * 1000.0, "p50_ms": stats[0], "p99_ms": stats[1], "max_ms": stats[2],
                     "missed_deadlines": int(stats[3]), "utilisation": stats[4] / 1000.0 / (blocks * period),
                     "state_GB_per_gpu": Cn * per_ch / 1e9,
                     "path": "fcb_fftconv_process, pinned host buffers, one synchronous call per block, back to back "
                             "(max over ranks of every figure)"}
        lib.fcb_host_free(p_in)
        lib.fcb_host_free(p_out)
        conv.close()
    return out


def run_mimo(local, rank, world, steps, warmup):
    """BASELINE configs[4] under the driver's eyes (N > 1): 16 x 16 matrix, 10 s IRs, block 512, the IR partitions
    sharded over the N GPUs; 1 and 16 streams (k_mac_rt, the register-tiled matrix MAC on the FP32 pipes) and 128 streams
    (tcgen05 K4); partial spectra exchanged by peer stores over NVLink and by NCCL, the peer form also with K3 overlapped
    with the next block's MAC (`_overlap`); block 520 is checked against the
    unsharded engine (a failed check is reported in the entry, it does not end the run: the ranks must stay in step)."""
    import torch
    import torch.distributed as dist
    import fft_convolution_b200 as F
    from fft_convolution_b200.distributed import ShardedMimoConvolver
    N, B, L = 16, 512, 10 * SAMPLE_RATE
    h = synth_irs(0, N * N, 0, L).reshape(N, N, L)
    res = {"config": f"MIMO {N}x{N}, IR 10 s ({L} taps, S = {(L + B - 1) // B}), block {B}, IR partitions sharded over {world} GPUs"}
    for NS in (1, 16, 128):
        x = [torch.from_numpy(synth_noise(0, NS * N, B * i, B)).cuda(local) for i in range(8)]
        ref = None
        NCHK = 520  # past half of the 938-slot ring: every shard's segment range has met real input spectra
        if rank == 0:  # the unsharded engine on the same blocks; the last one is compared
            try:  # rank 0 alone runs this: a failure here must not leave the other ranks alone in the collectives below
                whole = F.MimoConvolver.init(h, B, L, n_streams=NS, device=local)
                o = torch.empty((NS * N, B), dtype=torch.float32, device=f"cuda:{local}")
                for i in range(NCHK):
                    whole.partial_dev(x[i % 8].data_ptr(), B)
                    whole.finish_dev(o.data_ptr(), B)
                whole.sync()
                ref = o.cpu().numpy().copy()
                whole.close()
                del whole
            except Exception as exc:
                ref = None  # the entries of this stream count then carry parity_ok: null
                res[f"streams{NS}_unsharded_reference_error"] = f"{type(exc).__name__}: {exc}"[:300]
        R = NS * N
        # overlap: K3 of a block (the kernel that waits for the peers) beside K1 and the MAC of the next block
        modes = [("peer", False, False), ("nccl", False, False), ("peer", False, True)]
        if NS > 1 and R % world == 0:  # reduce-scatter: every rank finishes its own R / N output rows
            modes += [("peer", True, False), ("nccl", True, False), ("peer", True, True)]
        for exchange, scatter, overlap in modes:
            m = ShardedMimoConvolver(h, B, L, n_streams=NS, device=local, exchange=exchange, scatter=scatter, overlap=overlap)
            out = torch.empty((NS * N, B), dtype=torch.float32, device=f"cuda:{local}")
            err = None
            for i in range(NCHK):
                m.process_dev(x[i % 8], out)
            torch.cuda.synchronize()
            if ref is not None:
                lo, hi = m.rows  # every row, or this rank's rows in the reduce-scatter form
                got = out.cpu().numpy()[lo:hi]
                err = float(np.max(np.abs(got - ref[lo:hi]))) / max(float(np.sqrt(np.mean(ref[lo:hi].astype(np.float64) ** 2))), 1e-9)
                if err > 1e-5:
                    print(f"bench.py: sharded matrix ({NS} streams, {exchange}) differs from the unsharded engine: {err:.3e} x RMS", file=sys.stderr)
            for i in range(warmup):
                m.process_dev(x[i % 8], out)
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(m.stream)
            for i in range(steps):
                m.process_dev(x[i % 8], out)
            m.join()  # overlap: the last block's K3 belongs to the timed region
            e1.record(m.stream)
            torch.cuda.synchronize()
            m.m.sync()  # surfaces a peer-exchange timeout
            ms = reduce_max([e0.elapsed_time(e1) / steps], device=f"cuda:{local}")[0]
            res[f"streams{NS}_{exchange}" + ("_reduce_scatter" if scatter else "") + ("_overlap" if overlap else "")] = {
                "ms_per_block": ms, "realtime_factor": 1000.0 * B / SAMPLE_RATE / ms, "tensor_cores": bool(m.m.uses_tensor_cores),
                "mac_kernel": m.m.mac_kernel, "parity_ok": None if err is None else bool(err <= 1e-5),
                "output": "sharded by row over the ranks" if scatter else "complete on every rank",
                "blocks_in_flight": "K3 of block n beside the MAC of block n + 1 (throughput of queued blocks)" if overlap else "one at a time",
                "T_cmac_per_s": NS * N * N * ((L + B - 1) // B) * B / (ms / 1e3) / 1e12,
                "max_abs_err_over_rms_vs_unsharded_after_520_blocks": err}
            m.m.close()
            del m
            dist.barrier()
    return res


# ---------------------------------------------------------------------------------------------
def run_b200(args) -> None:
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator comes up; stdout must carry
        # the ONE JSON line only, so route fd 1 to stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    import fft_convolution_b200 as F
    from fft_convolution_b200 import _lib
    lib = _lib.load()
    if args.mac_impl is not None:
        _lib.check(lib.fcb_tune(b"mac_impl", args.mac_impl))
    if args.mac_stages is not None:
        _lib.check(lib.fcb_tune(b"mac_stages", args.mac_stages))
    if args.pipe_group is not None:
        _lib.check(lib.fcb_tune(b"pipe_group", args.pipe_group))
    fused = args.fused != 0 and args.block <= 512 and args.block >= 32
    _lib.check(lib.fcb_tune(b"fused_block", 1 if fused else 0))
    if args.tma_io is not None:
        _lib.check(lib.fcb_tune(b"tma_io", args.tma_io))
    if args.zero_copy is not None:
        _lib.check(lib.fcb_tune(b"zero_copy", args.zero_copy))
    if args.fused_stages is not None:
        _lib.check(lib.fcb_tune(b"fused_stages", args.fused_stages))
    if args.k1_late is not None:
        _lib.check(lib.fcb_tune(b"k1_late", args.k1_late))

    Cn, B = args.channels, args.block
    L = int(args.ir_seconds * SAMPLE_RATE)
    chan0, _ = rank_channel_range(rank, world, Cn)
    stream = torch.cuda.Stream(device=local)
    t0 = time.time()
    irs = synth_irs(chan0, Cn, 0, L)
    t_gen = time.time() - t0
    conv = F.FFTConvolver.init(irs, B, L, device=local, stream=stream.cuda_stream)
    S, K = conv.seg_count, conv.block_size + 1
    del irs

    # ---- correctness of the timed path: first blocks of a few channels vs f64 truth (numpy) ----
    NCHK = 6
    x_chk = synth_noise(chan0, Cn, 0, B * NCHK)
    y_chk = np.zeros((Cn, B * NCHK), np.float32)
    blk_out = np.zeros((Cn, B), np.float32)
    for b in range(NCHK):
        conv.process(np.ascontiguousarray(x_chk[:, b * B:(b + 1) * B]), blk_out)
        y_chk[:, b * B:(b + 1) * B] = blk_out
    worst = 0.0
    for c in sorted({0, Cn // 2, Cn - 1}):
        h = synth_irs(chan0 + c, 1, 0, L)[0].astype(np.float64)
        n = B * NCHK
        nfft = 1 << int(np.ceil(np.log2(2 * n)))
        truth = np.fft.irfft(np.fft.rfft(x_chk[c].astype(np.float64), nfft) * np.fft.rfft(h[:n], nfft), nfft)[:n]
        worst = max(worst, float(np.max(np.abs(y_chk[c] - truth)) / np.sqrt(np.mean(truth ** 2))))
    if worst > 1e-5:
        raise SystemExit(f"bench.py: parity check failed: max-abs err {worst:.3e} x RMS > 1e-5")

    # ---- device-resident inputs: 8 distinct blocks, rotated ------------------------------------
    NIN = 8
    with torch.cuda.stream(stream):
        d_in = [torch.from_numpy(synth_noise(chan0, Cn, B * (NCHK + i), B)).cuda(local, non_blocking=False) for i in range(NIN)]
        d_out = torch.empty((Cn, B), dtype=torch.float32, device=f"cuda:{local}")

    def step_dev(i):
        conv.process_dev(d_in[i % NIN].data_ptr(), B, B, d_out.data_ptr(), B, B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step_dev(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.fcb_launch_count()
    _lib.check(lib.fcb_profile_mac(1))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(args.steps):
        step_dev(args.warmup + i)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    tot_ms, nl = C.c_double(), C.c_uint64()
    _lib.check(lib.fcb_profile_mac_read(C.byref(tot_ms), C.byref(nl)))
    _lib.check(lib.fcb_profile_mac(0))
    launches = lib.fcb_launch_count() - launches0
    clocks = sampler.stop()

    # ---- the same step over a window of at least a second (the driver's K may be 20 steps = 18 ms) ----
    n_sus = max(args.steps, int(1200.0 / max(ms / args.steps, 1e-3)))
    barrier()
    ev0.record(stream)
    for i in range(n_sus):
        step_dev(i)
    ev1.record(stream)
    barrier()
    sus_ms = reduce_max([ev0.elapsed_time(ev1)], device=f"cuda:{local}")[0]
    sustained = {"steps": n_sus, "window_ms": sus_ms, "ms_per_step": sus_ms / n_sus,
                 "value": aggregate_value(world, Cn, n_sus, B, sus_ms), "unit": UNIT}
    launches += n_sus

    # ---- strong scaling: BASELINE configs[3] read as 4096 channels in total, sharded over the N GPUs ----
    strong = None
    if world > 1 and Cn % world == 0:
        Cs = Cn // world
        conv_s = F.FFTConvolver.init(synth_irs(rank * Cs, Cs, 0, L), B, L, device=local, stream=stream.cuda_stream)
        n_str = max(args.steps, 200)
        for i in range(args.warmup):
            conv_s.process_dev(d_in[i % NIN].data_ptr(), B, B, d_out.data_ptr(), B, B)
        barrier()
        ev0.record(stream)
        for i in range(n_str):
            conv_s.process_dev(d_in[i % NIN].data_ptr(), B, B, d_out.data_ptr(), B, B)
        ev1.record(stream)
        barrier()
        st_ms = reduce_max([ev0.elapsed_time(ev1)], device=f"cuda:{local}")[0]
        strong = {"channels_total": Cn, "channels_per_gpu": Cs, "steps": n_str, "ms_per_step": st_ms / n_str,
                  "value": Cn * n_str * B / SAMPLE_RATE / (st_ms / 1000.0), "unit": UNIT, "scaling": "strong"}
        launches += n_str + args.warmup
        conv_s.close()

    # ---- end to end through the host-pointer API: pinned H2D + K1..K3 + D2H every step ---------
    nbytes = Cn * B * 4
    p_in, p_out = lib.fcb_host_alloc(nbytes), lib.fcb_host_alloc(nbytes)
    if not p_in or not p_out:
        raise SystemExit("bench.py: pinned allocation failed")
    h_in = np.ctypeslib.as_array(C.cast(p_in, C.POINTER(C.c_float)), shape=(Cn, B))
    h_out = np.ctypeslib.as_array(C.cast(p_out, C.POINTER(C.c_float)), shape=(Cn, B))
    h_in[:] = synth_noise(chan0, Cn, B * (NCHK + NIN), B)
    for _ in range(args.warmup):
        conv.process(h_in, h_out)
    barrier()
    launches_e0 = lib.fcb_launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        conv.process(h_in, h_out)  # synchronous: returns with the block's output in host memory
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    launches += lib.fcb_launch_count() - launches_e0
    barrier()

    # ---- where the end-to-end overhead sits: the same synchronous call pattern on DEVICE buffers (launch + kernel + host
    # sync per step, no PCIe data): the difference to `ms_per_step` is the host's launch/sync cost, the rest of the e2e gap is
    # the data path
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_dev(i)
        conv.sync()
    sync_s = time.perf_counter() - t0
    launches += args.steps
    barrier()

    ms_max, e2e_ms_max, sync_ms_max = reduce_max([ms, e2e_s * 1000.0, sync_s * 1000.0], device=f"cuda:{local}")
    value = aggregate_value(world, Cn, args.steps, B, ms_max)
    e2e_value = aggregate_value(world, Cn, args.steps, B, e2e_ms_max)

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    # K2 (three-launch path): per channel 16*(S-1)*K read (IR rows + ring rows of segments 1..S-1;
    # SURVEY §8(d)'s 16*S*K with segment 0 moved into K3) + 8*K pre_multiplied written.
    # Fused kernel: the same MAC stream plus everything K1 and K3 touch — 8*K IR segment 0 read,
    # 8*K ring slot written, 4*B input read, 4*B overlap read, 4*B overlap written, 4*B output written.
    mac_bytes = 16 * (S - 1) * K
    if fused:
        bytes_per_launch = Cn * (mac_bytes + 8 * K + 8 * K + 16 * B)
        kname = "k_block_fused (K1+K2+K3 in one launch; K2's delay-line MAC is %.1f %% of its bytes)" % (
            100.0 * mac_bytes / (mac_bytes + 16 * K + 16 * B))
    else:
        bytes_per_launch = Cn * (mac_bytes + 8 * K)
        kname = "k_mac_bulk (K2: delay-line complex MAC)"
    k2_ms = tot_ms.value / max(nl.value, 1)
    achieved = bytes_per_launch / (k2_ms / 1000.0) / 1e9
    peak, peak_src = measured_peak_gbs()
    traffic = None
    tp = ROOT / "profiles" / ("fused_traffic.json" if fused else "k2_traffic.json")
    if tp.exists():
        try:
            t = json.loads(tp.read_text())
            # an ncu capture is only quoted for the kernels it was taken from (scripts/capture_traffic.sh stamps the hash)
            if t.get("channels") == Cn and t.get("kernel_source_sha256_16") == kernel_source_hash():
                traffic = t.get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "frac_of_nominal_8TBs": achieved / 8000.0, "bytes_per_launch": bytes_per_launch,
                "avg_launch_ms": k2_ms, "launches_timed": int(nl.value),
                "kernel_share_of_step": tot_ms.value / ms if ms > 0 else None}

    realtime = None
    if args.realtime:
        lib.fcb_host_free(p_in)
        lib.fcb_host_free(p_out)
        p_in = p_out = None
        conv.close()
        del d_in, d_out
        torch.cuda.empty_cache()
        try:  # an optional block must never cost the headline line
            realtime = run_realtime(F, lib, local, rank, world, chan0, Cn, e2e_ms_max / args.steps, B, L, args.realtime_blocks, barrier)
        except Exception as exc:  # e.g. another tenant holding HBM: report, do not die
            realtime = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    mimo = None
    if world > 1 and args.mimo:
        if conv._h:
            conv.close()
        torch.cuda.empty_cache()
        try:
            mimo = run_mimo(local, rank, world, 200, 20)
        except Exception as exc:
            mimo = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if rank == 0:
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            import oracle  # noqa: F401  (cpu_baseline leg only)
            threads = host_threads()
            # ~10 s of CPU work: 16 channels per host thread (each thread streams ~25 MB of state per
            # block, like a real many-channel host), 4 s of audio per channel
            ch, calls = min(max(threads * 16, 16), 512), 375
            v, secs = cpu_port_run(ch, B, L, calls, threads)
            cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": f"{ch} channels x {calls} blocks of {B} ({calls * B / SAMPLE_RATE:.1f} s of audio each), block loop only, "
                                      f"{secs:.2f} s wall on {threads} threads",
                            "note": "CPU restatement of the reference algorithm, not rustfft (no Rust toolchain)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_dict(args),
            "tuning": {"segments": S, "mac_impl": args.mac_impl, "mac_stages": args.mac_stages, "fused_block": bool(fused),
                       "ir_gen_s": round(t_gen, 1), "numa": numa},
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                    "ms_per_step": e2e_ms_max / args.steps,
                    "ms_per_step_same_calls_on_device_buffers": sync_ms_max / args.steps,
                    "path": "fcb_fftconv_process on pinned host buffers: the whole-block kernel pulls each channel's input "
                            "block from host memory and pushes its output block back with cp.async.bulk (one launch per step)"
                            if (fused and args.zero_copy != 0 and args.tma_io != 0) else
                            "fcb_fftconv_process on pinned host buffers: H2D / kernel / D2H pipelined over channel groups"},
            "gpu_launches": int(launches), "clocks": clocks,
            "parity": {"max_abs_err_over_rms_vs_f64": worst, "tolerance": 1e-5},
            "realtime_headroom": {"block_period_ms": 1000.0 * B / SAMPLE_RATE,
                                  "block_time_ms": ms_max / args.steps},
            "sustained": sustained, "strong": strong, "realtime": realtime, "mimo": mimo,
            "scaling_note": "headline = weak (4096 channels PER GPU); `strong` = 4096 channels in total over the N GPUs",
        }
        print(json.dumps(line), flush=True)
    if p_in:
        lib.fcb_host_free(p_in)
        lib.fcb_host_free(p_out)
    if conv._h:
        conv.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--channels", type=int, default=4096, help="channels per GPU")
    ap.add_argument("--block", type=int, default=512)
    ap.add_argument("--ir-seconds", type=float, default=2.0)
    ap.add_argument("--mac-impl", type=int, default=None)
    ap.add_argument("--mac-stages", type=int, default=None)
    ap.add_argument("--pipe-group", type=int, default=None)
    ap.add_argument("--fused-stages", type=int, default=None)
    ap.add_argument("--zero-copy", type=int, default=None)
    ap.add_argument("--tma-io", type=int, default=None)
    ap.add_argument("--k1-late", type=int, default=None, help="0: forward FFT before the MAC stream even when the block comes from host memory")
    ap.add_argument("--fused", type=int, default=1, help="1: one fused K1+K2+K3 kernel per block (default), 0: three launches")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--realtime", type=int, default=1, help="1: also run the measured real-time block (12 500 channels and the largest fitting count)")
    ap.add_argument("--realtime-blocks", type=int, default=1000)
    ap.add_argument("--mimo", type=int, default=1, help="1: with N > 1 also run BASELINE configs[4] sharded over the N GPUs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)  # timing rule: at least 3 warm-up steps
    args.steps = max(args.steps, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
