#!/usr/bin/env python
"""Per-launch device times of one fp16 infer (CUDA events between launches), median of a few repeats.
    python scripts/time_layers.py [1080p|4k|540p] [repeats] [fp16|int8-mixed]      (env knobs: HDRTV_RING_MAX, HDRTV_WAVES, HDRTV_MIN_BAND)"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import hdr_realtime_video_pipeline_b200 as hb  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "1080p"
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 5
h, w = {"540p": (540, 960), "1080p": (1080, 1920), "4k": (2160, 3840)}[wl]
prec = sys.argv[3] if len(sys.argv) > 3 else "fp16"
wfile = "tests/golden/weights_int8_mixed_qat.npz" if prec == "int8-mixed" else "tests/golden/weights_hr.npz"
net = hb.HDRTVNetB200(os.path.join(REPO, wfile), precision=prec, warmup_passes=0, use_hg=False)
frame = torch.from_numpy(hb.synth_frame(0, h, w)).cuda()
x, c = net.preprocess_device(frame)
for _ in range(2):
    net.infer((x, c))
runs = [net.time_plan((x, c)) for _ in range(rep)]
names = [n for n, _ in runs[0]]
ms = np.median(np.array([[t for _, t in r] for r in runs]), axis=0)
px = h * w
print(f"# {wl} env RING_MAX={os.environ.get('HDRTV_RING_MAX')} WAVES={os.environ.get('HDRTV_WAVES')} total={ms.sum():.3f} ms")
for n, t in zip(names, ms):
    print(f"{t * 1000:9.1f} us  {n}")
