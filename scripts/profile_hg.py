#!/usr/bin/env python
"""Small fixed workload for ncu: the HG stage (stage-in, 19 gconv launches, tail) on a fixed base output, one warm-up pass
+ N profiled passes.   python scripts/profile_hg.py [1080p|4k|540p] [passes]      (21 launches per pass)"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import hdr_realtime_video_pipeline_b200 as hb  # noqa: E402
from hdr_realtime_video_pipeline_b200.synth import hg_random_state_dict  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "1080p"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
h, w = {"540p": (540, 960), "1080p": (1080, 1920), "4k": (2160, 3840)}[wl]
net = hb.HDRTVNetB200(os.path.join(REPO, "tests/golden/weights_hr.npz"), precision="fp16", warmup_passes=0, use_hg=True,
                      hg_weights=hg_random_state_dict(0))
rng = np.random.default_rng(0)
base = torch.from_numpy((0.55 + 0.45 * rng.random((1, 3, h, w))).astype(np.float16)).cuda()
for _ in range(1 + n):
    net.hg_stage(base)
torch.cuda.synchronize()
print("done", wl, n)
