#!/usr/bin/env python
"""HG stage time against highlight coverage (the device-side highlight gate, DESIGN fact 20): everywhere / none / one spot /
a few random spots, at 1080p and 4K.   python scripts/hg_gate_probe.py"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np, torch
import hdr_realtime_video_pipeline_b200 as hb
from hdr_realtime_video_pipeline_b200.synth import hg_random_state_dict
net = hb.HDRTVNetB200(os.path.join(REPO, 'tests/golden/weights_hr.npz'), precision='fp16', warmup_passes=0, use_hg=True, hg_weights=hg_random_state_dict(0))
rng = np.random.default_rng(3)
for (h, w) in ((1080, 1920), (2160, 3840)):
    dark = (0.70 * rng.random((1, 3, h, w))).astype(np.float16)
    def spot(n):
        f = dark.copy()
        r = np.random.default_rng(n)
        for _ in range(n):
            y, x = int(r.integers(0, h - 20)), int(r.integers(0, w - 20))
            f[0, :, y:y + 16, x:x + 16] = np.float16(0.9)
        return f
    def ms(base):
        t = torch.from_numpy(base).cuda()
        for _ in range(3): net.hg_stage(t)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): net.hg_stage(t)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 10
    full = (0.55 + 0.45 * rng.random((1, 3, h, w))).astype(np.float16)
    one = dark.copy(); one[0, :, h // 2:h // 2 + 16, w // 2:w // 2 + 16] = np.float16(0.9)
    print(f"{w}x{h}: highlights everywhere {ms(full):.3f} ms | none {ms(dark):.3f} | one 16x16 highlight {ms(one):.3f} | 3 random {ms(spot(3)):.3f} | 10 random {ms(spot(10)):.3f}")
