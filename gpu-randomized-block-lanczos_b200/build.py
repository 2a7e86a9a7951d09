"""Builds librbl_b200.so in-tree with nvcc for sm_100a (no torch involved: the library is plain CUDA C++).

Every source is compiled to its own object (in parallel, only when it or a header changed), then linked."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "librbl_b200.so")
SOURCES = ["kernels.cu", "rowops.cu", "spmm.cu", "spmm_sched.cu", "spmm_lab.cu", "reorth_tc16.cu", "reorth_f64.cu", "microbench.cu", "solver.cu", "capi.cu",
           "band_eig.cpp", "comm.cpp", "partition.cpp", "mmio.cpp"]
SOURCES = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
HEADERS = [f for f in os.listdir(CSRC) if f.endswith(".h")] + [os.path.join("..", "..", "include", "rbl_b200.h")]
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC,-O3,-mavx2,-mfma"]


def _mtime(path: str) -> float:
    return os.path.getmtime(path) if os.path.exists(path) else 0.0


def _stale_objects(force: bool):
    hdr = max(_mtime(os.path.join(CSRC, h)) for h in HEADERS)
    out = []
    for s in SOURCES:
        obj = os.path.join(OBJDIR, os.path.splitext(s)[0] + ".o")
        if force or _mtime(obj) < max(_mtime(os.path.join(CSRC, s)), hdr, _mtime(os.path.abspath(__file__))):
            out.append((s, obj))
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    newest_input = max([_mtime(os.path.join(CSRC, f)) for f in SOURCES + HEADERS] + [_mtime(os.path.abspath(__file__))])
    if not force and os.path.exists(LIB) and _mtime(LIB) >= newest_input:
        return LIB      # up to date (the objects need not exist: they do not travel to the GPU box)
    todo = _stale_objects(force)
    objs = [os.path.join(OBJDIR, os.path.splitext(s)[0] + ".o") for s in SOURCES]

    def compile_one(item):
        src, obj = item
        cmd = [nvcc] + FLAGS + (["-Xptxas=-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        return src, r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        for src, r in ex.map(compile_one, todo):
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n" + r.stdout + r.stderr)
            if verbose:
                sys.stderr.write(r.stderr)
    r = subprocess.run([nvcc, "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-ldl", "-lpthread"],
                       cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
