"""Torch-CPU functional port of the reference's eager path (TEST / BASELINE INFRASTRUCTURE ONLY).

The reference's CPU configuration (scripts/setup_cpu.ps1: stock torch wheel, HDRTVNetTorch(device="cpu",
precision="fp32", compile_model=False, use_hg=False)) runs Ensemble_AGCM_LE through ATen/oneDNN.  The reference
tree does not exist on the GPU box, so `bench.py`'s cpu_baseline / `--impl reference` legs time THIS restatement:
the same graph, the same ATen ops, all host threads (kind = "port").  tests/test_oracle_golden.py pins it against
the numpy oracle and the reference-generated fixtures.  Never imported by the product package.

Graph follows Condition_arch.py:19-35, 559-585; HDRUNet3T1_arch.py:152-206; arch_util.py:60-95;
hdrtvnet_torch.py:2239-2296 (preprocess), :2352-2368 (postprocess).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def to_torch_state(sd: dict) -> dict:
    return {k: torch.from_numpy(np.ascontiguousarray(np.asarray(v, np.float32))) for k, v in sd.items()}


def preprocess(frame_bgr: np.ndarray):
    raw = torch.from_numpy(np.ascontiguousarray(frame_bgr)).flip(2).permute(2, 0, 1).unsqueeze(0)
    x = raw.to(torch.float32).mul_(1.0 / 255.0)
    cond = F.interpolate(x, scale_factor=0.25, mode="bicubic", align_corners=False, recompute_scale_factor=False,
                         antialias=True)
    return x, cond


def _conv(sd, name, x, stride=1):
    w = sd[name + ".weight"]
    return F.conv2d(x, w, sd[name + ".bias"], stride=stride, padding=w.shape[-1] // 2)


def _classifier(sd, cond):
    p = "AGCM.classifier.model."
    x = cond
    for ci, ni in ((0, 3), (4, 7), (8, 11), (12, 15), (16, None)):
        x = F.conv2d(x, sd[f"{p}{ci}.weight"], sd[f"{p}{ci}.bias"])
        x = F.avg_pool2d(x, 3, stride=2, padding=1, count_include_pad=True)
        x = F.leaky_relu(x, 0.2)
        if ni is not None:
            x = F.instance_norm(x, weight=sd[f"{p}{ni}.weight"], bias=sd[f"{p}{ni}.bias"], eps=1e-5)
    x = F.conv2d(x, sd[p + "20.weight"], sd[p + "20.bias"])
    return F.adaptive_avg_pool2d(x, 1).flatten(1)


def _agcm(sd, x, cond):
    fea = _classifier(sd, cond)

    def lin(n):
        return F.linear(fea, sd[f"AGCM.{n}.weight"], sd[f"AGCM.{n}.bias"]).view(1, -1, 1, 1)

    o = _conv(sd, "AGCM.conv_first", x)
    o = F.relu(o * lin("cond_scale_first") + lin("cond_shift_first") + o)
    o = _conv(sd, "AGCM.HRconv", o)
    o = F.relu(o * lin("cond_scale_HR") + lin("cond_shift_HR") + o)
    o = _conv(sd, "AGCM.conv_last", o)
    return o * lin("cond_scale_last") + lin("cond_shift_last") + o


def _sft(sd, p, fea, cond):
    scale = _conv(sd, p + ".SFT_scale_conv1", F.leaky_relu(_conv(sd, p + ".SFT_scale_conv0", cond), 0.1))
    shift = _conv(sd, p + ".SFT_shift_conv1", F.leaky_relu(_conv(sd, p + ".SFT_shift_conv0", cond), 0.1))
    return fea * (scale + 1) + shift


def _resblock(sd, p, x, cond):
    fea = _sft(sd, p + ".sft1", x, cond)
    fea = F.relu(_conv(sd, p + ".conv1", fea))
    fea = _sft(sd, p + ".sft2", fea, cond)
    return x + _conv(sd, p + ".conv2", fea)


def _seq(sd, p, x, spec):
    for idx, stride, act in spec:
        x = _conv(sd, f"{p}.{idx}", x, stride)
        if act:
            x = F.leaky_relu(x, 0.1)
    return x


def _align(x, ref):
    rh, rw = ref.shape[-2:]
    x = x[..., :rh, :rw] if (x.shape[-2] >= rh and x.shape[-1] >= rw) else x
    ph, pw = rh - x.shape[-2], rw - x.shape[-1]
    if ph > 0 or pw > 0:
        x = F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2), mode="replicate")
    return x


def _le(sd, img):
    p = "LE."
    cond = _seq(sd, p + "cond_first", img, [(0, 1, True), (2, 1, True), (4, 1, True)])
    cond1 = _seq(sd, p + "CondNet1", cond, [(0, 1, True), (2, 1, True), (4, 1, False)])
    cond2 = _seq(sd, p + "CondNet2", cond, [(0, 2, True), (2, 1, True), (4, 1, False)])
    cond3 = _seq(sd, p + "CondNet3", cond, [(0, 2, True), (2, 2, True), (4, 1, False)])
    cond4 = _seq(sd, p + "CondNet4", cond, [(0, 2, True), (2, 2, True), (4, 2, False)])
    fea0 = F.relu(_conv(sd, p + "conv_first", img))
    fea0 = _sft(sd, p + "SFT_layer1", fea0, cond1)
    fea0 = F.relu(_conv(sd, p + "HR_conv1", fea0))
    fea1 = _resblock(sd, p + "recon_trunk1.0", F.relu(_conv(sd, p + "down_conv1", fea0, 2)), cond2)
    fea2 = _resblock(sd, p + "recon_trunk2.0", F.relu(_conv(sd, p + "down_conv2", fea1, 2)), cond3)
    fea3 = F.relu(_conv(sd, p + "down_conv3", fea2, 2))
    out = fea3
    for i in range(4):
        out = _resblock(sd, f"{p}recon_trunk3.{i}", out, cond4)
    out = out + fea3
    up = _align(F.relu(F.pixel_shuffle(_conv(sd, p + "up_conv1.0", out), 2)), fea2)
    out = _resblock(sd, p + "recon_trunk4.0", up + fea2, cond3)
    up = _align(F.relu(F.pixel_shuffle(_conv(sd, p + "up_conv2.0", out), 2)), fea1)
    out = _resblock(sd, p + "recon_trunk5.0", up + fea1, cond2)
    up = _align(F.relu(F.pixel_shuffle(_conv(sd, p + "up_conv3.0", out), 2)), fea0)
    out = _sft(sd, p + "SFT_layer2", up + fea0, cond1)
    out = F.relu(_conv(sd, p + "HR_conv2", out))
    return img + _align(_conv(sd, p + "conv_last", out), img)


@torch.inference_mode()
def infer(sd: dict, x: torch.Tensor, cond: torch.Tensor):
    agcm_out = _agcm(sd, x, cond)
    return _le(sd, agcm_out), agcm_out


@torch.inference_mode()
def process(sd: dict, frame_bgr: np.ndarray) -> np.ndarray:
    """HDRTVNetTorch.process on CPU: BGR u8 -> BGR u8."""
    x, cond = preprocess(frame_bgr)
    out, _ = infer(sd, x, cond)
    t = out.squeeze(0).clamp_(0.0, 1.0).mul_(255.0).add_(0.5).to(torch.uint8).flip(0).permute(1, 2, 0).contiguous()
    return t.numpy()


@torch.inference_mode()
def process_rgb48(sd: dict, frame_bgr: np.ndarray) -> np.ndarray:
    """preprocess -> infer -> feeder pack (gui_pipeline_worker_feeders.py:237-249)."""
    x, cond = preprocess(frame_bgr)
    out, _ = infer(sd, x, cond)
    f = out.squeeze(0).permute(1, 2, 0).clamp(0.0, 1.0).contiguous().numpy().astype(np.float32)
    np.multiply(f, 65535.0, out=f)
    np.add(f, 0.5, out=f)
    np.clip(f, 0.0, 65535.0, out=f)
    return f.astype(np.uint16)
