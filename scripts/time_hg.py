#!/usr/bin/env python
"""Per-launch device times of the HG stage (FP16, tcgen05 gconv kernels) with achieved TFLOP/s per launch, the whole
AGCM+LE+HG frame, and - when baseline/_ref is present - the reference's own CUDA FP16 eager HG_Composite on the same GPU.

    python scripts/time_hg.py [1080p|4k|540p] [repeats] [--ref]
"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import hdr_realtime_video_pipeline_b200 as hb  # noqa: E402
from hdr_realtime_video_pipeline_b200.synth import hg_random_state_dict, hg_state_dict_spec  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "1080p"
rep = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 5
h, w = {"540p": (540, 960), "1080p": (1080, 1920), "4k": (2160, 3840)}[wl]
hp, wp = (h + 31) // 32 * 32, (w + 31) // 32 * 32
LEVEL = {"conv1.0": 0, "conv2.0": 1, "conv3_1.0": 1, "conv3_2.0": 2, "conv4_1.0": 2, "conv4_2.0": 3, "conv5_1.0": 3, "conv5_2.0": 4,
         "conv_code1.0": 4, "conv_code2.0": 5, "Up_conv1.0": 5, "conv6": 4, "Up_conv2.0": 4, "conv7": 3, "Up_conv3.0": 3, "conv8": 2,
         "Up_conv4.0": 2, "conv9": 1, "Up_conv5.0": 1}
spec = hg_state_dict_spec()


def flops(layer):
    o, c, k, _ = spec[layer + ".weight"]
    lv = LEVEL[layer]
    return 2.0 * o * c * k * k * (hp >> lv) * (wp >> lv)


hg_sd = hg_random_state_dict(0)
net = hb.HDRTVNetB200(os.path.join(REPO, "tests/golden/weights_hr.npz"), precision="fp16", warmup_passes=0, use_hg=True, hg_weights=hg_sd)
frame_np = hb.synth_frame(3, h, w, "white_salt")
frame = torch.from_numpy(frame_np).cuda()
x, c = net.preprocess_device(frame)
for _ in range(2):
    out, _ = net.infer((x, c))
base = net._gpu_out.clone()
runs = [net.hg_time_plan(base) for _ in range(rep)]
names = [n for n, _ in runs[0]]
ms = np.median(np.array([[t for _, t in r] for r in runs]), axis=0)
tot_f = 0.0
rows = []
print(f"# HG stage {wl} ({hp}x{wp} padded), per launch (median of {rep})")
for n, t in zip(names, ms):
    layer = n.split()[0][3:]
    f = flops(layer)
    tot_f += f
    rows.append({"launch": n, "ms": float(t), "tflops": f / (t * 1e-3) / 1e12})
    print(f"{t * 1000:9.1f} us  {f / (t * 1e-3) / 1e12:7.0f} TFLOP/s  {n}")
print(f"# sum {ms.sum():.3f} ms, {tot_f / 1e12:.3f} TFLOP per frame -> {tot_f / (ms.sum() * 1e-3) / 1e12:.0f} TFLOP/s over the gconv launches")


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


n = 20 if wl != "4k" else 10
t_stage = timed(lambda: net.hg_stage(base), n)
t_frame = timed(lambda: net.infer(net.preprocess_device(frame, assume_ready=True)), n)
res = {"workload": wl, "hg_stage_ms": t_stage, "frame_agcm_le_hg_ms": t_frame, "frames_per_s": 1000.0 / t_frame,
       "hg_stage_tflops": tot_f / (t_stage * 1e-3) / 1e12, "launches": rows}
print(f"# HG stage alone {t_stage:.3f} ms ({res['hg_stage_tflops']:.0f} TFLOP/s incl. stage-in / tail); AGCM+LE+HG frame {t_frame:.3f} ms = {1000 / t_frame:.1f} frames/s")
if "--ref" in sys.argv:
    from oracle import reference_loader as RL
    ref = RL.load()
    if ref is not None:
        import tempfile
        path = os.path.join(tempfile.mkdtemp(), "HG.pt")
        torch.save({k: torch.from_numpy(np.asarray(v)) for k, v in hg_sd.items()}, path)
        rnet = ref.HDRTVNetTorch(ref.weights("HR.pt"), device="cuda", precision="fp16", compile_model=False, use_hg=True, hg_weights=path,
                                 warmup_passes=0)
        with torch.inference_mode():
            xr, cr = rnet.preprocess(frame_np)
            t_ref = timed(lambda: rnet.infer((xr, cr)), max(3, n // 3))
        res["reference_cuda_fp16_eager_infer_ms"] = t_ref
        print(f"# reference (unmodified, CUDA FP16 eager, cudnn.benchmark) AGCM+LE+HG infer: {t_ref:.2f} ms = {1000 / t_ref:.1f} frames/s")
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
with open(os.path.join(REPO, "gpurun_out", f"hg_time_{wl}.json"), "w") as f:
    json.dump(res, f, indent=1)
