"""In-tree build of the engine with nvcc for sm_100a (cross-compiles without a GPU).

    python -m hdr_realtime_video_pipeline_b200.build [--force] [-v]

Two shared libraries from the same sources:
    libhdrtv_b200.so        the product: the C ABI of include/hdrtv_b200.h and nothing else
    libhdrtv_b200_test.so   test build (-DHDRTV_TEST_EXPORTS): the same ABI plus the debug / self-test / micro-probe entry
                            points of include/hdrtv_b200_test.h (tests/ and scripts/ only)
"""
from __future__ import annotations

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT = os.path.join(_HERE, "libhdrtv_b200.so")
OUT_TEST = os.path.join(_HERE, "libhdrtv_b200_test.so")
SOURCES = ["engine.cu"]
HEADERS = ["common.cuh", "ptx.cuh", "conv_p8.cuh", "chain_p8.cuh", "conv2x_p8.cuh", "conv3z_pair.cuh", "hg.cuh", "hg_engine.cuh", "letterbox_engine.cuh", "probes.cuh", "kernels_f32.cuh", "kernels_io.cuh",
           os.path.join("..", "..", "include", "hdrtv_b200.h"), os.path.join("..", "..", "include", "hdrtv_b200_test.h")]
NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-std=c++17"]


def _stale(out: str) -> bool:
    if not os.path.isfile(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS if os.path.isfile(os.path.join(CSRC, f)))


def build(force: bool = False, verbose: bool = False, test_lib: bool = True) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("HDRTV_NVCC_EXTRA", "").split()          # e.g. -DHDRTV_CHAIN_TRACE for scripts/chain_trace.py
    jobs = []
    for out, defs in ((OUT, []), (OUT_TEST, ["-DHDRTV_TEST_EXPORTS"])):
        if out == OUT_TEST and not test_lib:
            continue
        if not force and not _stale(out):
            continue
        cmd = [nvcc, *NVCC_FLAGS, *defs, *extra, *(["-Xptxas", "-v"] if verbose else []), *SOURCES, "-o", out]
        jobs.append((out, subprocess.Popen(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for out, proc in jobs:                                          # the two builds run side by side
        so, se = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed for {os.path.basename(out)}:\n" + so + se)
        if verbose:
            print(se)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
