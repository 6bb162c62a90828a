"""Compile csrc/ into libpsiformer_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m psiformer_torch_b200.build [--force]

The library is built IN-TREE (psiformer_torch_b200/csrc/libpsiformer_b200.so) so that it travels
with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(CSRC, "libpsiformer_b200.so")
STAMP = LIB + ".srchash"
SOURCES = ["psif_api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-lcuda",
]


def _source_hash() -> str:
    h = hashlib.sha256()
    files = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    files = [os.path.join(CSRC, f) for f in files] + [os.path.join(ROOT, "include", "psiformer_b200.h")]
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == _source_hash()


def build(force: bool = False, verbose: bool = False, stats: bool = False) -> str:
    """``stats``: compile the GEMM's wait-cycle counters in (tools/ss_stats.py); such a build is never 'current'."""
    if not force and not stats and is_current():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    if stats:
        cmd.insert(1, "-DPSIF_SS_STATS=1")
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building libpsiformer_b200.so")
    if verbose:
        sys.stderr.write(proc.stderr)
    with open(STAMP, "w") as fh:
        fh.write("stats build" if stats else _source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, stats="--stats" in sys.argv))
