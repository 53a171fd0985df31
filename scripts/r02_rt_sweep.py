#!/usr/bin/env python
"""k_mac_rt knob sweep in one process (16 x 16 matrix, 10 s IRs, block 512): python scripts/r02_rt_sweep.py "16:mimo_rt_r=2,mimo_rt_waves=3" "32:" ...
One JSON line per spec: whole-block and MAC-kernel time."""
import ctypes as C
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from fft_convolution_b200 import _lib, MimoConvolver  # noqa: E402

DEFAULTS = {"mimo_rt": 1, "mimo_rt_wb": 1, "mimo_rt_r": 4, "mimo_rt_waves": 0, "mimo_rt_min": 1, "mimo_tc": 0}


def main():
    lib = _lib.load()
    N, B, L = 16, 512, 480000
    h = bench.synth_irs(0, N * N, 0, L).reshape(N, N, L)
    steps = int(os.environ.get("STEPS", 40))
    for spec in sys.argv[1:]:
        ns, _, tunes = spec.partition(":")
        NS = int(ns)
        kv = dict(DEFAULTS)
        kv.update({k: int(v) for k, v in (t.split("=") for t in filter(None, tunes.split(",")))})
        for k, v in kv.items():
            _lib.check(lib.fcb_tune(k.encode(), v))
        m = MimoConvolver.init(h, B, L, n_streams=NS, tensor_cores=None if kv["mimo_tc"] == 2 else bool(kv["mimo_tc"]))
        x = [torch.from_numpy(bench.synth_noise(0, NS * N, B * i, B)).cuda() for i in range(4)]
        out = torch.empty((NS * N, B), dtype=torch.float32, device="cuda")
        st = torch.cuda.ExternalStream(lib.fcb_mimo_stream(m._h))

        def block(i):
            m.partial_dev(x[i % 4].data_ptr(), B)
            m.finish_dev(out.data_ptr(), B)

        for i in range(5):
            block(i)
        torch.cuda.synchronize()
        _lib.check(lib.fcb_profile_mac(1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(steps):
            block(i)
        e1.record(st)
        torch.cuda.synchronize()
        tot, nl = C.c_double(), C.c_uint64()
        _lib.check(lib.fcb_profile_mac_read(C.byref(tot), C.byref(nl)))
        _lib.check(lib.fcb_profile_mac(0))
        print(json.dumps({"streams": NS, "mac_kernel": m.mac_kernel, "tune": {k: v for k, v in kv.items() if DEFAULTS[k] != v},
                          "ms_per_block": round(e0.elapsed_time(e1) / steps, 4), "mac_ms": round(tot.value / max(nl.value, 1), 4)}), flush=True)
        m.close()
        del m, x, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
