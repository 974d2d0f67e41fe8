#!/usr/bin/env python
"""Fixtures for the reference's shipping INT8 layout on NVIDIA, "INT8 Mixed QAT" (gui_config.py:162;
configs/qat_layouts/original_nohg_mixed_w8a8.txt: 29 W8A8 / 78 W8A16 / 21 FP16 layers), made by RUNNING THE REFERENCE's own
eager INT8 model (HDRTVNetTorch(precision="int8-mixed"), hdrtvnet_torch.py:296-410, 1748-1963) on CPU in the build container.

    weights_int8_mixed_qat.npz      the raw checkpoint arrays (int8 weights, scales, activation quantisers)
    int8mixed_<class>_<HxW>.npz     frame, out, agcm_out of the reference run
    int8mixed_layers_64x96.npz      quantised input q (uint8), output and the EXACT integer accumulators sum(q * w_int8)
                                    (what a kind::i8 MMA must reproduce bit for bit, computed here in float64/int64) of W8A8
                                    modules of every kind, recorded with forward hooks

    python scripts/make_golden_int8_mixed.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import reference_loader as RL  # noqa: E402
from hdr_realtime_video_pipeline_b200.synth import synth_frame  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


def main():
    ref = RL.load(allow_source_tree=True)
    torch.set_num_threads(os.cpu_count() or 8)
    path = ref.weights("HR_original_int8_mixed_qat.pt")
    ck = torch.load(path, map_location="cpu", weights_only=False)
    raw = {}
    for k, v in ck["state_dict"].items():
        a = v.detach().cpu()
        raw[k] = a.numpy() if a.dtype == torch.int8 else a.float().numpy()
    np.savez_compressed(os.path.join(OUT, "weights_int8_mixed_qat.npz"), **raw)
    proc = ref.HDRTVNetTorch(path, device="cpu", precision="int8-mixed", compile_model=False, use_hg=False, warmup_passes=0,
                             predequantize="off")
    for cls, h, w, idx in (("noise", 64, 96, 0), ("ramps", 72, 100, 1), ("noise", 136, 248, 4)):
        frame = synth_frame(idx, h, w, cls)
        with torch.inference_mode():
            t, cnd = proc.preprocess(frame)
            res = proc.infer((t.clone(), cnd.clone()))
        np.savez_compressed(os.path.join(OUT, f"int8mixed_{cls}_{h}x{w}.npz"), frame=frame, out=res[0].float().numpy().copy(),
                            agcm_out=res[1].float().numpy().copy())
        print(f"  int8mixed_{cls}_{h}x{w}.npz")
    layers = ["LE.down_conv1", "LE.CondNet3.0", "LE.CondNet4.4", "LE.up_conv1.0", "LE.recon_trunk3.0.conv1", "LE.CondNet2.4"]
    rec = {}
    mods = dict(proc.model.named_modules())
    hooks = []
    for n in layers:
        def mk(name):
            def f(mod, inp, out):
                x = inp[0].detach().float()
                rec[name + "|out"] = out.detach().float().numpy().copy()
                # the layer sees its input only through the quantiser (W8A8Conv2d.forward :350-360): keep q (uint8) instead of
                # the fp32 input, and the exact integer accumulators sum(q * w_int8) of this module on it
                q = ((x - mod.x_zero) / mod.x_scale).round().clamp(0, 255).to(torch.float64)
                acc = F.conv2d(q, mod.weight_int8.to(torch.float64), None, mod.stride, mod.padding, mod.dilation, mod.groups)
                rec[name + "|acc"] = acc.numpy().astype(np.int32)
                rec[name + "|q"] = q.numpy().astype(np.uint8)
                rec[name + "|stride"] = np.array(mod.stride[0])
            return f
        hooks.append(mods[n].register_forward_hook(mk(n)))
    frame = synth_frame(0, 64, 96, "noise")
    with torch.inference_mode():
        t, cnd = proc.preprocess(frame)
        proc.infer((t.clone(), cnd.clone()))
    for hk in hooks:
        hk.remove()
    rec["layers"] = np.array(layers)
    np.savez_compressed(os.path.join(OUT, "int8mixed_layers_64x96.npz"), **rec)
    print("  int8mixed_layers_64x96.npz", os.path.getsize(os.path.join(OUT, "int8mixed_layers_64x96.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
