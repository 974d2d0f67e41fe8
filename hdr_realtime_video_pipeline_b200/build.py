"""In-tree build of libhdrtv_b200.so with nvcc for sm_100a (cross-compiles without a GPU).

    python -m hdr_realtime_video_pipeline_b200.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT = os.path.join(_HERE, "libhdrtv_b200.so")
SOURCES = ["engine.cu"]
HEADERS = ["common.cuh", "ptx.cuh", "conv_p8.cuh", "chain_p8.cuh", "conv2x_p8.cuh", "probes.cuh", "kernels_f32.cuh", "kernels_io.cuh",
           os.path.join("..", "..", "include", "hdrtv_b200.h")]
NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-std=c++17"]


def _stale() -> bool:
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("HDRTV_NVCC_EXTRA", "").split()          # e.g. -DHDRTV_CHAIN_TRACE for scripts/chain_trace.py
    cmd = [nvcc, *NVCC_FLAGS, *extra, *(["-Xptxas", "-v"] if verbose else []), *SOURCES, "-o", OUT]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
