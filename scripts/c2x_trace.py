#!/usr/bin/env python
"""Timeline of the two-conv kernel (CTA 0): per row and role, clock64 stamps.  Needs a library built with
HDRTV_NVCC_EXTRA=-DHDRTV_CHAIN_TRACE.
    python scripts/c2x_trace.py [1080p|4k] [launch-name-substring]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import hdr_realtime_video_pipeline_b200 as hb  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "1080p"
which = sys.argv[2] if len(sys.argv) > 2 else "trunk1.0"
h, w = {"540p": (540, 960), "1080p": (1080, 1920), "4k": (2160, 3840)}[wl]
net = hb.HDRTVNetB200(os.path.join(REPO, "tests/golden/weights_hr.npz"), precision="fp16", warmup_passes=0, use_hg=False, debug_library=True)
frame = torch.from_numpy(hb.synth_frame(0, h, w)).cuda()
x, c = net.preprocess_device(frame)
net.infer((x, c))
torch.cuda.synchronize()
names = [n for n, _ in net.time_plan((x, c))]
first = next(i for i, n in enumerate(names) if n.startswith("LE."))
idx = next(i for i, n in enumerate(names) if which in n and "+" in n) - first
print("#", names[idx + first], "plan index", idx)
tr = net.chain_trace(agcm=False, index=idx)
t0 = tr[tr > 0].min()
roles = {0: "prod  [top, slot free]", 1: "mmaA  [top, tempty, in_full, issued]",
         2: "mmaB  [top, tempty, mid_full, issued]", 3: "epiA  [top, st_full, S loaded, a_tfull, acc loaded, math, mid_empty, stored]",
         4: "epiB  [top, res_full, b_tfull, acc loaded, stored]"}
for r, d in roles.items():
    print("# role", r, d)
for row in range(20, 40):
    for role in range(5):
        vals = [int(v - t0) if v > 0 else None for v in tr[row, role]]
        out, prev = [], None
        for v in vals:
            out.append("      -" if v is None else (f"{v:7d}" if prev is None else f"{v - prev:+7d}"))
            prev = v if v is not None else prev
        print(f"{row:3d} {role} | " + " ".join(out))
