"""Builds csrc/liba3d.so (the C-ABI library of include/a3d.h) in-tree with nvcc
for sm_100a.  ``python -m articulation3d_b200.build``"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", "a3d.cu")]
OUT = os.path.join(HERE, "csrc", "liba3d.so")

NVCC_FLAGS = [
    "-shared", "-Xcompiler", "-fPIC", "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-fmad=false",                       # belt and braces: the exact paths use *_rn intrinsics
    "-Xcompiler", "-ffp-contract=off",   # host helpers: every float64 operation rounded separately
    "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64",
    "-I", os.path.join(ROOT, "include"),
]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build(force: bool = False, verbose: bool = False) -> str:
    deps = SRC + [os.path.join(ROOT, "include", "a3d.h"), os.path.abspath(__file__)]
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in deps)):
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRC
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if verbose:
        sys.stderr.write(proc.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
