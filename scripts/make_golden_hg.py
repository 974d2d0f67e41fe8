#!/usr/bin/env python
"""Fixtures for the HG stage, made by RUNNING THE REFERENCE in the build container (CPU).

HG.pt is absent from the reference tree (.MISSING_LARGE_BLOBS), so the highlight generator gets the seeded stand-in
weights of ``hdr_realtime_video_pipeline_b200.synth.hg_random_state_dict`` (regenerated from the seed wherever they are
needed: 37 M parameters do not belong in git); the base model is the shipped HR.pt.  What runs is the reference's own
``HG_Composite`` (HG_Composite_arch.py) - base model, mask, reflect pad, Hallucination_Generator with real eval-mode
BatchNorm, crop - in FP32 and, for the half-precision triangle, ``.half()``.

    python scripts/make_golden_hg.py            -> tests/golden/hg_*.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
REF = os.environ.get("HDRTV_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "src"))

from hdr_realtime_video_pipeline_b200.synth import hg_random_state_dict, synth_frame  # noqa: E402
from models.hdrtvnet_modules.HG_Composite_arch import HG_Composite  # noqa: E402
from oracle import hdrtvnet_oracle as O  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")
CASES = [("noise", 64, 96, 0), ("white_salt", 72, 100, 3), ("mixed", 96, 160, 1)]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    base_sd = {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(os.path.join(GOLD, "weights_hr.npz")).items()}
    hg_sd = hg_random_state_dict(0)
    model = HG_Composite(classifier="color_condition", cond_c=6, in_nc=3, out_nc=3, nf=32, act_type="relu",
                         weighting_network=False, hg_nf=64, mask_r=0.75).eval()
    model.base.load_state_dict(base_sd, strict=True)
    model.hg.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in hg_sd.items()}, strict=True)
    half = None
    for cls, h, w, idx in CASES:
        if cls == "mixed":                     # left half ramps, right half white + salt: a mask edge inside the frame
            frame = synth_frame(idx, h, w, "ramps")
            frame[:, w // 2:] = synth_frame(idx, h, w, "white_salt")[:, w // 2:]
        else:
            frame = synth_frame(idx, h, w, cls)
        x, cond = O.preprocess(frame, np.float32)
        with torch.inference_mode():
            xt, ct = torch.from_numpy(x), torch.from_numpy(cond)
            base_out, _ = model.base((xt, ct))
            hg_out, cond_out = model((xt, ct))
            mask = model._make_mask(base_out, r=0.75)
            rec = dict(frame=frame, base_out=base_out.numpy(), hg_out=hg_out.numpy(), mask=mask.numpy().astype(np.uint8),
                       agcm_out=cond_out.numpy())
            try:                                  # half-precision leg (CPU half convs exist in torch >= 2.2)
                if half is None:
                    import copy
                    half = copy.deepcopy(model).half()
                o16, _ = half((xt.half(), ct.half()))
                b16, _ = half.base((xt.half(), ct.half()))
                rec["hg_out_fp16"] = o16.float().numpy()
                rec["hg_out_fp16_dtype"] = np.asarray(str(o16.dtype))
                rec["base_out_fp16"] = b16.float().numpy()
            except Exception as exc:              # pragma: no cover
                print("half leg skipped:", exc)
        name = f"hg_{cls}_{h}x{w}.npz"
        np.savez_compressed(os.path.join(GOLD, name), **rec)
        print(name, "mask on", float(mask.mean()), "max|hg-base|", float((hg_out - base_out).abs().max()),
              "out dtype fp16 leg", rec.get("hg_out_fp16_dtype"))


if __name__ == "__main__":
    main()
