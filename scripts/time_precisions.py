#!/usr/bin/env python
"""ms/frame of the non-default precisions (FP32 CUDA-core parity path, INT8 Full-QAT layout on that path) next to FP16.
    python scripts/time_precisions.py [1080p|4k|540p] [frames]"""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import hdr_realtime_video_pipeline_b200 as hb  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "1080p"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
h, w = {"540p": (540, 960), "1080p": (1080, 1920), "4k": (2160, 3840)}[wl]
G = os.path.join(REPO, "tests", "golden")
for precision, weights in (("fp16", "weights_hr.npz"), ("fp32", "weights_hr.npz"), ("int8-full", "weights_int8_full_qat.npz")):
    net = hb.HDRTVNetB200(os.path.join(G, weights), precision=precision, warmup_passes=0, use_hg=False)
    frames = [torch.from_numpy(hb.synth_frame(i, h, w)).cuda() for i in range(2)]
    for i in range(2):
        net.infer(net.preprocess_device(frames[i % 2]))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        net.infer(net.preprocess_device(frames[i % 2]))
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / n * 1e3
    print(f"{wl} {precision:9s}: {ms:9.2f} ms/frame ({1e3 / ms:7.1f} frames/s), workspace {net.workspace_bytes() / 2**30:.2f} GiB", flush=True)
    net.close()
    del net
    torch.cuda.empty_cache()
