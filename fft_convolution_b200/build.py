"""Build libfftconv_b200.so in-tree with nvcc for sm_100a (no other architecture, no fallback)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
SO = PKG / "libfftconv_b200.so"
SOURCES = ["engine.cu", "host_mirror.cu", "mimo.cu"]
HEADERS = ["common.cuh", "engine_internal.cuh", "fft_kernels.cuh", "mac_kernels.cuh", "fused_kernel.cuh", "mimo_tc.cuh", "mimo_rt.cuh", "offline_kernels.cuh", "../../include/fftconv_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=default",
    "--cudart", "static",
]


# extra nvcc flags for A/B builds, e.g. FCB_NVCC_EXTRA="-DFCB_FFT_E16_FROM=12"
NVCC_FLAGS += [f for f in os.environ.get("FCB_NVCC_EXTRA", "").split() if f]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libfftconv_b200.so cannot be built")


def stale() -> bool:
    if not SO.exists():
        return True
    t = SO.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [(CSRC / h).resolve() for h in HEADERS] + [Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not stale():
        return SO
    nvcc = nvcc_path()
    objdir = PKG / "_build"
    objdir.mkdir(exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = objdir / (Path(src).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [nvcc, "-shared", "--cudart", "static", "-o", str(SO), *objs]
    subprocess.run(cmd, check=True)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
