#!/usr/bin/env python
"""Reference-run fixtures at the BASELINE config sizes (960x540, 1920x1080, 3840x2160) for the FP16 parity gate.

Runs the reference's own ``Ensemble_AGCM_LE`` (HDRUNet3T1 + AGCM; Ensemble_AGCM_LE_arch.py:889-897) on CPU in FP32 and
as ``model.half()`` (the reference's FP16 path, hdrtvnet_torch.py:2164-2167) on deterministic synthetic frames and keeps
a SAMPLE of the outputs (the whole frames would be 12-50 MB each):

    sub32 / sub16      out[:, ::S, ::S]                      S = 8 (<= 1080p) or 16 (4K)
    rows32 / rows16    rows {0,1,2, H/2-2..H/2+1, H-3,H-2,H-1} in full
    cols32 / cols16    columns {0,1,2, 124..132, 250..260, W/2-1, W/2, W-3,W-2,W-1} in full  (strip seams of the 126-
                       and 128-pixel tiles of the CUDA kernels)
    agcm32 / agcm16    agcm_out[:, ::S, ::S]
    stats              mean / min / max of the full FP32 and FP16 outputs
    dref               max |ref16 - ref32| over the FULL frame (the reference's own FP16 error)

``tests/test_gpu_parity.py::test_fp16_config_sizes_match_reference`` compares the CUDA path on the same frames at the
same positions.  The frames are regenerated from (class, index) by ``synth_frame``.

    python scripts/make_golden_large.py [--only 1080]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import reference_loader as RL  # noqa: E402
from oracle import hdrtvnet_oracle as O  # noqa: E402  (random_state_dict only)
from hdr_realtime_video_pipeline_b200.synth import synth_frame  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
CASES = [  # (tag, weights, class, frame index, H, W)
    ("net540", "hr", "noise", 0, 540, 960), ("net540", "rand0", "noise", 0, 540, 960), ("net540", "hr", "ramps", 1, 540, 960),
    ("net1080", "hr", "noise", 0, 1080, 1920), ("net1080", "hr", "ramps", 1, 1080, 1920), ("net1080", "rand0", "noise", 0, 1080, 1920),
    ("net1080", "hr", "white_salt", 3, 1080, 1920),
    ("net2160", "hr", "noise", 0, 2160, 3840), ("net2160", "hr", "ramps", 1, 2160, 3840),
]


def sample_positions(h: int, w: int):
    s = 16 if h > 1080 else 8
    rows = sorted({0, 1, 2, h // 2 - 2, h // 2 - 1, h // 2, h // 2 + 1, h - 3, h - 2, h - 1})
    cols = sorted({0, 1, 2, *range(124, 133), *range(250, 261), w // 2 - 1, w // 2, w - 3, w - 2, w - 1})
    return s, np.array(rows), np.array([c for c in cols if c < w])


def ref_preprocess(frame, dtype):
    """hdrtvnet_torch.py:2252-2294 (CPU branch).  ATen has no CPU Half kernel for the antialiased resize; its CUDA kernel
    accumulates half inputs in fp32 and rounds once, which is what this does."""
    raw = torch.from_numpy(frame).flip(2).permute(2, 0, 1).unsqueeze(0)
    x = raw.to(dtype=dtype).mul_(1.0 / 255.0)
    cond = F.interpolate(x.float(), scale_factor=0.25, mode="bicubic", align_corners=False,
                         recompute_scale_factor=False, antialias=True).to(dtype)
    return x, cond


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    ref = RL.load(allow_source_tree=True)
    if ref is None:
        raise SystemExit("reference tree not found")
    torch.set_num_threads(os.cpu_count() or 8)
    hr = torch.load(ref.weights("HR.pt"), map_location="cpu", weights_only=True)
    sets = {"hr": {k: v.float() for k, v in hr.items()},
            "rand0": {k: torch.from_numpy(np.asarray(v)) for k, v in O.random_state_dict(0).items()}}
    models = {}
    for name, sd in sets.items():
        m = ref.Ensemble_AGCM_LE(classifier="color_condition", cond_c=6, in_nc=3, out_nc=3, nf=32, act_type="relu",
                                 weighting_network=False)
        m.load_state_dict(sd, strict=True)
        models[name] = m.eval().float()
    for tag, wname, cls, idx, h, w in CASES:
        if args.only and args.only not in tag:
            continue
        t0 = time.time()
        frame = synth_frame(idx, h, w, cls)
        m32 = models[wname]
        with torch.inference_mode():
            x, cond = ref_preprocess(frame, torch.float32)
            out32, agcm32 = m32((x, cond))
            out32, agcm32 = out32.numpy()[0].copy(), agcm32.numpy()[0].copy()
            m16 = m32.half()
            x16, cond16 = ref_preprocess(frame, torch.float16)
            out16, agcm16 = m16((x16, cond16))
            out16, agcm16 = out16.float().numpy()[0].copy(), agcm16.float().numpy()[0].copy()
            m16.float()
            m32.load_state_dict(sets[wname], strict=True)        # undo the fp16 rounding of the shared module
        s, rows, cols = sample_positions(h, w)
        name = f"{tag}_{wname}_{cls}.npz"
        np.savez_compressed(
            os.path.join(OUT, name), cls=np.array(cls), idx=np.array(idx), hw=np.array([h, w]), step=np.array(s), rows=rows, cols=cols,
            sub32=out32[:, ::s, ::s], sub16=out16[:, ::s, ::s].astype(np.float16),
            rows32=out32[:, rows, :], rows16=out16[:, rows, :].astype(np.float16),
            cols32=out32[:, :, cols], cols16=out16[:, :, cols].astype(np.float16),
            agcm32=agcm32[:, ::s, ::s], agcm16=agcm16[:, ::s, ::s].astype(np.float16),
            stats=np.array([out32.mean(), out32.min(), out32.max(), out16.mean(), out16.min(), out16.max()], np.float64),
            dref=np.array([np.abs(out16 - out32).max(), np.abs(out16 - out32).mean()], np.float64))
        print(f"  {name}: {os.path.getsize(os.path.join(OUT, name)) / 1024:.0f} KiB, dref max {np.abs(out16 - out32).max():.2e} "
              f"mean {np.abs(out16 - out32).mean():.2e}, {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main()
