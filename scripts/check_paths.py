#!/usr/bin/env python
"""Full-frame comparison of the frame paths at a bench resolution: every RGB48 frame produced by (a) process_rgb48
pipelined, (b) process_rgb48 serial, (c) preprocess -> infer -> tensor_to_rgb48_bytes, back to back without host
synchronisation, against frames produced one at a time with a device synchronisation after every call.
    python scripts/check_paths.py [1080p|4k|540p] [frames]"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import hdr_realtime_video_pipeline_b200 as hb  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "1080p"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
h, w = {"540p": (540, 960), "1080p": (1080, 1920), "4k": (2160, 3840)}[wl]
nd = int(sys.argv[3]) if len(sys.argv) > 3 else 8
net = hb.HDRTVNetB200(os.path.join(REPO, "tests/golden/weights_hr.npz"), precision="fp16", warmup_passes=0, use_hg=False)
frames = [torch.from_numpy(hb.synth_frame(i, h, w)).pin_memory() for i in range(nd)]
fnp = [f.numpy() for f in frames]
state = {}
want = []
for i in range(nd):
    torch.cuda.synchronize()
    fr = net.process_rgb48(fnp[i], serial=True)
    want.append(fr.numpy().copy())
    fr.release()
    torch.cuda.synchronize()
# the synchronised three-call path must agree with it
for i in range(nd):
    torch.cuda.synchronize()
    out = net.infer(net.preprocess(fnp[i]))
    torch.cuda.synchronize()
    fr = hb.tensor_to_rgb48_bytes(out, state)
    assert np.array_equal(fr.numpy(), want[i]), ("synchronised three-call path differs", i)
    fr.release()


def run(name, submit, in_flight):
    bad, pending = [], []

    def drain():
        j, fr = pending.pop(0)
        got = fr.numpy()
        if not np.array_equal(got, want[j % nd]):
            d = np.nonzero((got != want[j % nd]).any(axis=2))
            bad.append((j, int(d[0].size), int(d[0].min()), int(d[0].max()), int(d[1].min()), int(d[1].max())))
        fr.release()

    for i in range(n):
        pending.append((i, submit(i)))
        if len(pending) >= in_flight:
            drain()
    while pending:
        drain()
    print(f"{wl} {name}: {len(bad)} of {n} frames differ", bad[:6], flush=True)


run("one-call pipelined", lambda i: net.process_rgb48(fnp[i % nd]), 3)
run("one-call serial", lambda i: net.process_rgb48(fnp[i % nd], serial=True), 3)
run("three-call", lambda i: hb.tensor_to_rgb48_bytes(net.infer(net.preprocess(fnp[i % nd])), state), 2)
run("one-call pipelined (again)", lambda i: net.process_rgb48(fnp[i % nd]), 3)
