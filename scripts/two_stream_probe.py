#!/usr/bin/env python
"""Throughput of N independent contexts on N streams of one GPU (frames are independent: SURVEY §8e) against one context.
    python scripts/two_stream_probe.py [1080p|4k] [n_contexts] [frames] [--hg]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import hdr_realtime_video_pipeline_b200 as hb  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "1080p"
nctx = int(sys.argv[2]) if len(sys.argv) > 2 else 2
K = int(sys.argv[3]) if len(sys.argv) > 3 else 120
HG = "--hg" in sys.argv
h, w = {"540p": (540, 960), "1080p": (1080, 1920), "4k": (2160, 3840)}[wl]
W = os.path.join(REPO, "tests/golden/weights_hr.npz")
frames = [torch.from_numpy(hb.synth_frame(i, h, w)).cuda() for i in range(8)]


def run(n):
    if HG:
        from hdr_realtime_video_pipeline_b200.synth import hg_random_state_dict
        sd = hg_random_state_dict(0)
        nets = [hb.HDRTVNetB200(W, precision="fp16", warmup_passes=0, use_hg=True, hg_weights=sd) for _ in range(n)]
    else:
        nets = [hb.HDRTVNetB200(W, precision="fp16", warmup_passes=0, use_hg=False) for _ in range(n)]
    packers = [hb.RGB48Packer("cuda", ring_frames=3) for _ in range(n)]
    streams = [torch.cuda.Stream() for _ in range(n)]
    outs = [torch.empty((h, w, 3), dtype=torch.uint16, device="cuda") for _ in range(n)]

    def step(i):
        j = i % n
        with torch.cuda.stream(streams[j]):
            out = nets[j].infer(nets[j].preprocess_device(frames[i % 8], assume_ready=True))
            packers[j].pack_device(out, outs[j])

    for i in range(4 * n):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_event(e0)
    for i in range(K):
        step(i)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    for nn in nets:
        nn.close()
    return ms


for n in sorted({1, nctx}):
    ms = run(n)
    print(f"{wl}: {n} context(s): {ms:.3f} ms/frame = {1000 / ms:.1f} frames/s")
