"""Builds libvirolution_gpu.so (sm_100a only) in-tree with nvcc.  `python -m virolution_b200.build`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvirolution_gpu.so")
SOURCES = ["virolution_gpu.cu"]
HEADERS = ["vl_kernels.cuh", "vl_migrate.cuh", "vl_exchange.cuh", "vl_fused.cuh", "vl_rng.cuh", "vl_scan.cuh", "vl_state.cuh",
           os.path.join("..", "..", "include", "virolution_gpu.h")]

# -fmad=false: the fitness arithmetic must round like the reference's scalar f64 code (no FMA contraction of
# `1. + factor * f`, src/core/fitness/utility.rs:51-58); the path is HBM-bound so it costs nothing.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or stale():
        cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
        subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
