#!/usr/bin/env python
"""Install the UNMODIFIED reference next to the repo as ``baseline/_ref`` (git-ignored, travels to the GPU box).

The reference is a Windows PyQt application without setup.py / pyproject (``pip install /root/reference`` has nothing
to build), but its model wrapper, network modules, feeder pack and frame-processing mixin are plain Python that import
on Linux.  This recipe copies, byte for byte and without editing anything:

    src/*.py, src/models/**/*.py                         the wrapper (HDRTVNetTorch), the networks, feeders, frame processing
    src/models/weights/original/HR.pt                    the FP32/FP16 checkpoint
    src/models/weights/original/pytorch_int8/hr/*_qat.pt the Full-QAT and Mixed-QAT INT8 checkpoints (BASELINE config 5)
    configs/qat_layouts/*.txt                            the mixed W8A8 layouts

into ``baseline/_ref/`` with the same relative paths.  Nothing under ``baseline/_ref`` is tracked by git and nothing in
the product package imports it; it is used by

    bench.py --impl reference          the reference's own CPU eager path (setup_cpu), ``cpu_baseline.kind = "reference"``
    bench.py  gpu_eager_baseline       the reference's CUDA FP16 eager path on the same box (the bar to beat)
    tests/test_gpu_reference_live.py   FP16 parity at 1080p / 4K against the reference's CUDA FP16 output
    scripts/make_golden.py             (build container) fixtures

    python scripts/install_reference.py [--src /root/reference] [--force]
"""
from __future__ import annotations

import argparse
import glob
import hashlib
import json
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(REPO, "baseline", "_ref")

WEIGHTS = [
    "src/models/weights/original/HR.pt",
    "src/models/weights/original/pytorch_int8/hr/HR_original_int8_full_qat.pt",
    "src/models/weights/original/pytorch_int8/hr/HR_original_int8_mixed_qat.pt",
]


def _files(src: str) -> list[str]:
    rel = []
    for pat in ("src/*.py", "src/models/*.py", "src/models/hdrtvnet_modules/*.py", "configs/qat_layouts/*.txt"):
        rel += [os.path.relpath(p, src) for p in sorted(glob.glob(os.path.join(src, pat)))]
    rel += [w for w in WEIGHTS if os.path.isfile(os.path.join(src, w))]
    return rel


def install(src: str = "/root/reference", force: bool = False) -> str | None:
    """Returns the install directory, or None when the reference tree is not present (GPU box: uses the copy that
    travelled with the snapshot)."""
    if not os.path.isdir(os.path.join(src, "src", "models")):
        return DEST if os.path.isfile(os.path.join(DEST, "MANIFEST.json")) else None
    manifest_path = os.path.join(DEST, "MANIFEST.json")
    if os.path.isfile(manifest_path) and not force:
        return DEST
    files = _files(src)
    manifest = {}
    for rel in files:
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(manifest_path, "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    return DEST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=os.environ.get("HDRTV_REFERENCE", "/root/reference"))
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    d = install(a.src, a.force)
    if d is None:
        print("reference tree not found and no installed copy present", file=sys.stderr)
        sys.exit(1)
    n = len(json.load(open(os.path.join(d, "MANIFEST.json")))["files"])
    print(f"reference installed: {d} ({n} files)")
