#!/usr/bin/env python
"""Small fixed workload for ncu: 3 warm-up frames + N profiled frames of the device-resident hot path.
    python scripts/profile_frame.py [1080p|4k|540p] [frames]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import hdr_realtime_video_pipeline_b200 as hb  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "1080p"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
h, w = {"540p": (540, 960), "1080p": (1080, 1920), "4k": (2160, 3840)}[wl]
net = hb.HDRTVNetB200(os.path.join(REPO, "tests/golden/weights_hr.npz"), precision="fp16", warmup_passes=0, use_hg=False)
packer = hb.RGB48Packer("cuda")
frames = [torch.from_numpy(hb.synth_frame(i, h, w)).cuda() for i in range(4)]
out_dev = torch.empty((h, w, 3), dtype=torch.uint16, device="cuda")
l0 = net.launch_count() + packer.launch_count()
for i in range(3 + n):
    if i == 3:
        torch.cuda.synchronize()
        per = (net.launch_count() + packer.launch_count() - l0) // 3
        print("launches per frame:", per, flush=True)
    x, c = net.preprocess_device(frames[i % 4], assume_ready=True)
    packer.pack_device(net.infer((x, c)), out_dev)
torch.cuda.synchronize()
print("done", wl, n)
