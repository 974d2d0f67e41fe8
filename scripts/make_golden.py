#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (CPU) in the build
container.  The reference tree (/root/reference, read-only) does not exist on the
GPU box, so its outputs travel as these fixtures; this script is the committed
recipe that made them.

    python scripts/make_golden.py            # writes tests/golden/

What is pinned (reference file:line in each section):
  weights_hr.npz        HR.pt state-dict as fp32 arrays (src/models/weights/original/HR.pt)
  pre_*.npz             HDRTVNetTorch.preprocess (hdrtvnet_torch.py:2239-2296), fp32 + fp16 arithmetic
  aa_*.npz              F.interpolate(bicubic, antialias) on odd sizes (border taps)
  net_<w>_<case>.npz    Ensemble_AGCM_LE forward (Ensemble_AGCM_LE_arch.py:889-897): fea6, agcm_out,
                        out (fp32), out_fp16 (model.half() on CPU), postprocess BGR24, feeder RGB48
  net540_<w>.npz        same at 960x540 (BASELINE config 1), sub-sampled + statistics
  pack.npz              _tensor_to_rgb48_bytes CPU branch (gui_pipeline_worker_feeders.py:237-249) and
                        postprocess (hdrtvnet_torch.py:2352-2368) on edge-value tensors, fp32 + fp16
  pq.npz                _linear_bgr_to_bt2100_pq_bgr_u16 (gui_objective_metrics.py:531-539)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("HDRTV_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REF, "src"))

# gui_pipeline_worker_feeders imports the PyQt6 mpv widget; stub it.
_stub = types.ModuleType("gui_mpv_widget")
_stub.MpvHDRWidget = type("MpvHDRWidget", (), {})
sys.modules.setdefault("gui_mpv_widget", _stub)

from models.hdrtvnet_torch import HDRTVNetTorch  # noqa: E402
from models.hdrtvnet_modules.Ensemble_AGCM_LE_arch import Ensemble_AGCM_LE  # noqa: E402
import gui_pipeline_worker_feeders as feeders  # noqa: E402
import gui_objective_metrics as gom  # noqa: E402

from hdr_realtime_video_pipeline_b200.synth import synth_frame  # noqa: E402
from oracle import hdrtvnet_oracle as O  # noqa: E402  (only for random_state_dict)

OUT = os.path.join(REPO, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(8)
torch.manual_seed(0)


def save(name, **arrs):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrs)
    print(f"  {name}: {os.path.getsize(path) / 1024:.0f} KiB")


def build_model(sd_np):
    m = Ensemble_AGCM_LE(classifier="color_condition", cond_c=6, in_nc=3, out_nc=3, nf=32,
                         act_type="relu", weighting_network=False)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}, strict=True)
    return m.eval().float()


def ref_preprocess(frame, dtype):
    """hdrtvnet_torch.py:2252-2294 with the CPU branch, dtype arithmetic."""
    raw = torch.from_numpy(frame).flip(2).permute(2, 0, 1).unsqueeze(0)
    x = raw.to(dtype=dtype).mul_(1.0 / 255.0)
    # ATen has no CPU Half kernel for the antialiased resize; its CUDA kernel accumulates
    # half inputs in fp32 (accscalar_t) and rounds the result once, which is what this does.
    cond = F.interpolate(x.float(), scale_factor=0.25, mode="bicubic", align_corners=False,
                         recompute_scale_factor=False, antialias=True).to(dtype)
    return x, cond


def main():
    # ---- weights ---------------------------------------------------------
    hr_path = os.path.join(REF, "src/models/weights/original/HR.pt")
    hr = torch.load(hr_path, map_location="cpu", weights_only=True)
    hr_np = {k: v.float().numpy() for k, v in hr.items()}
    save("weights_hr.npz", **hr_np)
    weight_sets = {"hr": hr_np, "rand0": O.random_state_dict(0)}

    # ---- preprocess through the real wrapper + dtype variants ------------
    proc = HDRTVNetTorch(hr_path, device="cpu", precision="fp32", compile_model=False,
                         use_hg=False, warmup_passes=0)
    for (h, w) in ((64, 96), (72, 100), (135, 241)):
        frame = synth_frame(0, h, w, "noise")
        x, cond = proc.preprocess(frame)
        x16, cond16 = ref_preprocess(frame, torch.float16)
        save(f"pre_{h}x{w}.npz", frame=frame, x=x.numpy().copy(), cond=cond.numpy().copy(),
             x16=x16.numpy(), cond16=cond16.numpy())

    # ---- fast_condition_resize (bilinear condition image) through the real wrapper (hdrtvnet_torch.py:2268-2275)
    proc_fast = HDRTVNetTorch(hr_path, device="cpu", precision="fp32", compile_model=False, use_hg=False,
                              warmup_passes=0, fast_condition_resize=True)
    bil = {}
    for (h, w) in ((64, 96), (73, 101)):
        frame = synth_frame(0, h, w, "noise")
        x, cond = proc_fast.preprocess(frame)
        bil[f"frame_{h}x{w}"] = frame
        bil[f"cond_{h}x{w}"] = cond.numpy().copy()
        x16 = torch.from_numpy(frame).flip(2).permute(2, 0, 1).unsqueeze(0).to(torch.float16).mul_(1.0 / 255.0)
        bil[f"cond16_{h}x{w}"] = F.interpolate(x16.float(), scale_factor=0.25, mode="bilinear", align_corners=False,
                                               recompute_scale_factor=False).to(torch.float16).numpy()
    save("pre_bilinear.npz", **bil)
    if os.environ.get("HDRTV_GOLDEN_ONLY") == "bilinear":
        return

    # ---- INT8 Full-QAT layout (BASELINE config 5): the reference's eager INT8 model on CPU (fp32 compute fallback,
    # hdrtvnet_torch.py:1766-1773) on the shipped checkpoint; the raw checkpoint arrays travel as a fixture.
    int8_path = os.path.join(REF, "src/models/weights/original/pytorch_int8/hr/HR_original_int8_full_qat.pt")
    ck = torch.load(int8_path, map_location="cpu", weights_only=False)
    raw = {}
    for k, v in ck["state_dict"].items():
        a = v.detach().cpu()
        raw[k] = a.numpy() if a.dtype == torch.int8 else a.float().numpy()
    save("weights_int8_full_qat.npz", **raw)
    proc8 = HDRTVNetTorch(int8_path, device="cpu", precision="int8-full", compile_model=False, use_hg=False,
                          warmup_passes=0, predequantize="off")
    for cls, h, w, idx in (("noise", 64, 96, 0), ("ramps", 72, 100, 1), ("white_salt", 72, 100, 3)):
        frame = synth_frame(idx, h, w, cls)
        with torch.inference_mode():
            t, cnd = proc8.preprocess(frame)
            res = proc8.infer((t.clone(), cnd.clone()))
        out, agcm_out = res[0], res[1]
        save(f"int8_{cls}_{h}x{w}.npz", frame=frame, out=out.float().numpy().copy(), agcm_out=agcm_out.float().numpy().copy())
    # Fake-quantised networks amplify fp32 summation-order noise chaotically (a single flipped quantisation bucket
    # spreads through the U-Net), so end-to-end INT8 parity can only be statistical.  The exact pin is per layer:
    # (input, output) pairs recorded from the reference's own W8A8 modules on a small frame.
    layers = ["LE.conv_first", "LE.down_conv2", "LE.CondNet4.4", "LE.recon_trunk3.0.conv1", "LE.recon_trunk3.0.sft1.SFT_scale_conv0",
              "LE.recon_trunk3.0.sft1.SFT_shift_conv1", "LE.up_conv1.0", "AGCM.cond_scale_first", "AGCM.classifier.model.16"]
    rec = {}
    mods = dict(proc8.model.named_modules())
    hooks = []
    for n in layers:
        def mk(name):
            def f(mod, inp, out):
                rec[name + "|in"] = inp[0].detach().float().numpy().copy()
                o = out.detach().float().numpy().copy()
                rec[name + "|out"] = o[..., ::4, ::4] if (o.ndim == 4 and o.size > 60000) else o    # big maps: every 4th pixel
            return f
        hooks.append(mods[n].register_forward_hook(mk(n)))
    frame = synth_frame(0, 64, 96, "noise")
    with torch.inference_mode():
        t, cnd = proc8.preprocess(frame)
        proc8.infer((t.clone(), cnd.clone()))
    for hk in hooks:
        hk.remove()
    rec["layers"] = np.array(layers)
    save("int8_layers_64x96.npz", **rec)
    if os.environ.get("HDRTV_GOLDEN_ONLY") == "int8":
        return

    # ---- AA bicubic on awkward sizes -------------------------------------
    rng = np.random.default_rng(7)
    aa = {}
    for (h, w) in ((17, 23), (8, 9), (33, 64), (540 // 4, 31)):
        t = torch.from_numpy(rng.random((1, 3, h, w), dtype=np.float32))
        aa[f"in_{h}x{w}"] = t.numpy()
        aa[f"out_{h}x{w}"] = F.interpolate(t, scale_factor=0.25, mode="bicubic", align_corners=False,
                                           recompute_scale_factor=False, antialias=True).numpy()
    save("aa.npz", **aa)

    # ---- network, small cases --------------------------------------------
    cases = [("noise", 64, 96, 0), ("ramps", 72, 100, 1), ("black", 64, 96, 2), ("white_salt", 72, 100, 3),
             ("noise", 136, 248, 4)]
    for wname, sd_np in weight_sets.items():
        model = build_model(sd_np)
        model16 = build_model(sd_np).half()
        for cls, h, w, idx in cases:
            frame = synth_frame(idx, h, w, cls)
            with torch.inference_mode():
                x, cond = ref_preprocess(frame, torch.float32)
                fea = model.AGCM.classifier(cond).reshape(-1)
                out, agcm_out = model((x, cond))
                x16, cond16 = ref_preprocess(frame, torch.float16)
                out16, agcm16 = model16((x16, cond16))
                proc.model = model
                bgr24 = proc.postprocess(out.clone()).copy()
                rgb48 = np.frombuffer(feeders._tensor_to_rgb48_bytes(out.clone(), {}), dtype=np.uint16)
                rgb48 = rgb48.reshape(h, w, 3)
            save(f"net_{wname}_{cls}_{h}x{w}.npz", frame=frame, fea=fea.numpy(), agcm_out=agcm_out.numpy(),
                 out=out.numpy(), out_fp16=out16.float().numpy().astype(np.float16),
                 agcm_fp16=agcm16.float().numpy().astype(np.float16), bgr24=bgr24, rgb48=rgb48)

        # ---- BASELINE config 1 size: 960x540, subsampled --------------------
        h, w = 540, 960
        frame = synth_frame(0, h, w, "noise")
        with torch.inference_mode():
            x, cond = ref_preprocess(frame, torch.float32)
            fea = model.AGCM.classifier(cond).reshape(-1)
            out, agcm_out = model((x, cond))
            rgb48 = np.frombuffer(feeders._tensor_to_rgb48_bytes(out.clone(), {}), dtype=np.uint16).reshape(h, w, 3)
        o = out.numpy()
        save(f"net540_{wname}.npz", fea=fea.numpy(), out_sub=o[:, :, ::8, ::8].copy(),
             agcm_sub=agcm_out.numpy()[:, :, ::8, ::8].copy(),
             out_last_rows=o[:, :, -3:, :].copy(), out_stats=np.array([o.mean(), o.min(), o.max()], np.float64),
             rgb48_sub=rgb48[::8, ::8].copy(), rgb48_mean=np.array([rgb48.astype(np.float64).mean()]))

    # ---- pack edge values --------------------------------------------------
    rng = np.random.default_rng(11)
    vals = np.concatenate([
        np.array([-1.0, -0.0, 0.0, 1e-8, 0.5 / 65535, 1.0 / 65535, 0.25, 0.5, 0.75, 0.9996, 0.9997, 0.99999,
                  1.0, 1.0000001, 2.0, 65504.0], np.float32),
        rng.random(3 * 16 * 24 - 16, dtype=np.float32) * 1.2 - 0.1])
    t32 = torch.from_numpy(vals.reshape(1, 3, 16, 24).copy())
    t16 = t32.half()
    pack = {"in32": t32.numpy(), "in16": t16.numpy()}
    for tag, t in (("32", t32), ("16", t16)):
        pack[f"rgb48_{tag}"] = np.frombuffer(feeders._tensor_to_rgb48_bytes(t.clone(), {}), dtype=np.uint16).reshape(16, 24, 3)
        pack[f"bgr24_{tag}"] = proc.postprocess(t.clone()).copy()
    save("pack.npz", **pack)

    # ---- optional PQ transfer ----------------------------------------------
    lin = np.minimum(rng.random((16, 24, 3), dtype=np.float32) * 1.1 - 0.05, np.float32(1.0))  # peak<=1: no rescale (:81-83)
    lin[0, 0] = (0.0, 1.0, 0.5)
    pq_bgr = gom._linear_bgr_to_bt2100_pq_bgr_u16(lin[:, :, ::-1].copy(), peak_nits=1000.0)
    save("pq.npz", linear_rgb=lin, pq_rgb_u16=np.ascontiguousarray(pq_bgr[:, :, ::-1]))


if __name__ == "__main__":
    main()
