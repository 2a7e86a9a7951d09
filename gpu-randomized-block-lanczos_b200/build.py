"""Builds librbl_b200.so in-tree with nvcc for sm_100a (no torch involved: the library is plain CUDA C++)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "librbl_b200.so")
SOURCES = ["kernels.cu", "reorth_tc.cu", "reorth_tc16.cu", "reorth_f64.cu", "microbench.cu", "solver.cu", "capi.cu", "band_eig.cpp", "comm.cpp", "partition.cpp"]
HEADERS = ["kernels.h", "solver.h", "band_eig.h", "comm.h", "partition.h", os.path.join("..", "..", "include", "rbl_b200.h")]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [nvcc, "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC,-O3,-mavx2,-mfma", "--shared", "-o", LIB] + SOURCES + ["-ldl", "-lpthread"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
