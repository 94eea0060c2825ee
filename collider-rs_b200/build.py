"""Build libcollider_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

FMA contraction is disabled (-fmad=false): the f64 contract of DESIGN.md requires every product and sum
to be rounded separately, exactly like the reference's x86-64 SSE2 code.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "engine.cu")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("engine.cu", "kernels.cuh", "common.cuh", "narrowphase.cuh", "sort.cuh", "single.cuh", "tile.cuh", "drain.cuh", "migrate.cuh", "sort_cut.cuh", "collider.inl")] + [
    os.path.join(ROOT, "include", "collider_b200.h")
]
OUT = os.path.join(HERE, "lib", "libcollider_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def nvcc() -> str:
    path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(path):
        raise RuntimeError("nvcc not found")
    return path


def is_stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
