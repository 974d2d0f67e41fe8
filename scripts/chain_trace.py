#!/usr/bin/env python
"""Timeline of the fused layer-chain kernel (CTA 0): per step and row slot, clock64 stamps relative to the first.
    python scripts/chain_trace.py [1080p|4k] [agcm|cond]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import hdr_realtime_video_pipeline_b200 as hb  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "1080p"
which = sys.argv[2] if len(sys.argv) > 2 else "cond"
h, w = {"540p": (540, 960), "1080p": (1080, 1920), "4k": (2160, 3840)}[wl]
net = hb.HDRTVNetB200(os.path.join(REPO, "tests/golden/weights_hr.npz"), precision="fp16", warmup_passes=0, use_hg=False, debug_library=True)
frame = torch.from_numpy(hb.synth_frame(0, h, w)).cuda()
x, c = net.preprocess_device(frame)
net.infer((x, c))
torch.cuda.synchronize()
tr = net.chain_trace(agcm=(which == "agcm"), index=0)
t0 = tr[tr > 0].min()
print("# step slot | loop-top, issued, woke (tfull), tmem loaded, math done, tile stored, fenced, epilogue done  (deltas)")
for e in range(12, 26):
    for g in range(6):
        r = tr[e, g]
        vals = [int(v - t0) if v > 0 else None for v in r]
        out, prev = [], None
        for v in vals:
            out.append("      -" if v is None else (f"{v:7d}" if prev is None else f"{v - prev:+7d}"))
            prev = v if v is not None else prev
        print(f"{e:3d} {g} | " + " ".join(out))
