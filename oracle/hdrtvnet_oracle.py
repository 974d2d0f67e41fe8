"""CPU oracle for the HDRTVNet++ per-frame SDR->HDR path (TEST INFRASTRUCTURE ONLY).

This file is a plain-numpy restatement of the reference's algorithm for the hot
path.  It is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The shipped package
(``hdr_realtime_video_pipeline_b200``) never imports anything from ``oracle/``.

Parity pinning: the reference ships NO golden vectors or known-answer tests of
its own.  The oracle is therefore pinned against outputs of
the reference itself, run in the build container by
``scripts/make_golden.py`` (imports ``/root/reference/src`` read-only) and
committed under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every
function below against those fixtures.

Third-party arithmetic: conv / pool / instance-norm / interpolate live in
PyTorch ATen (reference pins torch==2.9.1; fixtures were generated with the
image's torch 2.11.0 CPU).  Their published definitions are restated here.

Each function cites the reference file:line it follows (paths relative to the
reference root).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

# --------------------------------------------------------------------------
# P1  preprocess  (src/models/hdrtvnet_torch.py:2239-2296)
# --------------------------------------------------------------------------

_INV255 = F32(1.0 / 255.0)  # python double 1/255 -> fp32 scalar, as torch's mul_(1.0/255.0)


def normalize_bgr_u8(frame_bgr: np.ndarray, dtype=np.float32) -> np.ndarray:
    """uint8 HxWx3 BGR -> (3,H,W) RGB, ``float(u8) * fp32(1/255)`` rounded to dtype.

    hdrtvnet_torch.py:2256-2261: flip(2) -> permute -> .to(dtype).mul_(1/255).
    For fp16 the reference converts u8->half (exact) and multiplies in half:
    the product of a half and the scalar is computed in fp32 and rounded once.
    """
    rgb = frame_bgr[:, :, ::-1].transpose(2, 0, 1)
    if np.dtype(dtype) == np.float16:
        # torch half mul_ with a python scalar: opmath is float, scalar kept in fp32.
        return (rgb.astype(F32) * _INV255).astype(np.float16)
    return rgb.astype(F32) * _INV255


def _cubic_aa_filter(x: np.ndarray) -> np.ndarray:
    """Keys cubic, a = -0.5 (ATen UpSampleKernel.cpp `aa_filter` for bicubic)."""
    a = F32(-0.5)
    x = np.abs(x).astype(F32)
    w1 = ((a + F32(2.0)) * x - (a + F32(3.0))) * x * x + F32(1.0)
    w2 = (((x - F32(5.0)) * x + F32(8.0)) * x - F32(4.0)) * a
    return np.where(x < 1.0, w1, np.where(x < 2.0, w2, F32(0.0))).astype(F32)


def aa_bicubic_weights(n_in: int, scale_factor: float = 0.25):
    """Per-output-index tap start, count and normalised weights of
    F.interpolate(mode='bicubic', antialias=True, align_corners=False,
    recompute_scale_factor=False) along one axis.  hdrtvnet_torch.py:2277-2285;
    ATen `_compute_indices_weights_aa`: scale = 1/scale_factor, support =
    2*scale, center = scale*(i+0.5), taps [int(center-support+0.5),
    int(center+support+0.5)) clipped to [0,n_in), weight k((j-center+0.5)/scale)
    normalised by the tap sum.
    """
    n_out = int(np.floor(n_in * scale_factor))
    scale = F32(1.0 / scale_factor)
    support = F32(2.0) * scale
    starts, counts, weights = [], [], []
    for i in range(n_out):
        center = scale * (F32(i) + F32(0.5))
        xmin = max(int(center - support + F32(0.5)), 0)
        xmax = min(int(center + support + F32(0.5)), n_in)
        j = np.arange(xmin, xmax, dtype=np.float32)
        w = _cubic_aa_filter((j - center + F32(0.5)) / scale)
        w = (w / w.sum(dtype=F32)).astype(F32)
        starts.append(xmin)
        counts.append(xmax - xmin)
        weights.append(w)
    return n_out, starts, counts, weights


def _aa_matrix(n_in: int, scale_factor: float = 0.25) -> np.ndarray:
    n_out, starts, counts, weights = aa_bicubic_weights(n_in, scale_factor)
    m = np.zeros((n_out, n_in), dtype=F32)
    for i in range(n_out):
        m[i, starts[i]:starts[i] + counts[i]] = weights[i]
    return m


def cond_downsample(x_chw: np.ndarray, dtype=np.float32) -> np.ndarray:
    """(3,H,W) -> (3,H//4,W//4) antialiased bicubic, fp32 accumulate, rounded to dtype."""
    c, h, w = x_chw.shape
    my = _aa_matrix(h)
    mx = _aa_matrix(w)
    xf = x_chw.astype(F32)
    # ATen CPU runs the horizontal pass first, then the vertical one; the CUDA kernel
    # (the fp16 path) accumulates the 2-D tap sum in fp32 and rounds once at the end.
    tmp = np.einsum("chw,ow->cho", xf, mx, optimize=True).astype(F32)
    out = np.einsum("chw,oh->cow", tmp, my, optimize=True).astype(F32)
    return out.astype(dtype)


def cond_bilinear(x: np.ndarray, dtype=np.float32) -> np.ndarray:
    """fast_condition_resize (hdrtvnet_torch.py:2268-2275): F.interpolate(scale_factor=0.25, mode="bilinear",
    align_corners=False, recompute_scale_factor=False).  Source coordinate (i+0.5)*4-0.5 = 4i+1.5: the mean of input
    pixels 4i+1 and 4i+2 in both directions (ATen accumulates half inputs in fp32 and rounds once)."""
    xf = x.astype(F32)
    c, h, w = xf.shape
    ho, wo = h // 4, w // 4
    iy1, iy2 = 4 * np.arange(ho) + 1, np.minimum(4 * np.arange(ho) + 2, h - 1)
    ix1, ix2 = 4 * np.arange(wo) + 1, np.minimum(4 * np.arange(wo) + 2, w - 1)
    top = F32(0.5) * xf[:, iy1][:, :, ix1] + F32(0.5) * xf[:, iy1][:, :, ix2]
    bot = F32(0.5) * xf[:, iy2][:, :, ix1] + F32(0.5) * xf[:, iy2][:, :, ix2]
    return (F32(0.5) * top + F32(0.5) * bot).astype(dtype)


def preprocess(frame_bgr: np.ndarray, dtype=np.float32, cond_mode: str = "bicubic_aa"):
    """Returns (x (1,3,H,W), cond (1,3,H//4,W//4)) like HDRTVNetTorch.preprocess; cond_mode "bilinear" is the
    reference's fast_condition_resize option, "zero" its HDRTVNET_ZERO_COND shortcut."""
    x = normalize_bgr_u8(frame_bgr, dtype)
    if cond_mode == "bilinear":
        cond = cond_bilinear(x, dtype)
    elif cond_mode == "zero":
        cond = np.zeros((3, x.shape[1] // 4, x.shape[2] // 4), dtype)
    else:
        cond = cond_downsample(x, dtype)
    return x[None], cond[None]


# --------------------------------------------------------------------------
# Primitive ops (ATen definitions restated)
# --------------------------------------------------------------------------

def conv2d(x: np.ndarray, w: np.ndarray, b: np.ndarray | None, stride: int = 1, pad: int | None = None) -> np.ndarray:
    """x (C,H,W), w (O,C,kh,kw) -> (O,Ho,Wo).  Zero padding, cross-correlation
    (torch.nn.Conv2d).  Tap-by-tap matmul accumulation in fp32."""
    o, c, kh, kw = w.shape
    if pad is None:
        pad = kh // 2
    _, h, wd = x.shape
    ho = (h + 2 * pad - kh) // stride + 1
    wo = (wd + 2 * pad - kw) // stride + 1
    xp = np.pad(x.astype(F32), ((0, 0), (pad, pad), (pad, pad))) if pad else x.astype(F32)
    out = np.zeros((o, ho * wo), dtype=F32)
    for ky in range(kh):
        for kx in range(kw):
            patch = xp[:, ky:ky + (ho - 1) * stride + 1:stride, kx:kx + (wo - 1) * stride + 1:stride]
            out += w[:, :, ky, kx].astype(F32) @ np.ascontiguousarray(patch).reshape(c, ho * wo)
    if b is not None:
        out += b.astype(F32)[:, None]
    return out.reshape(o, ho, wo)


def leaky_relu(x: np.ndarray, slope: float) -> np.ndarray:
    return np.where(x >= 0, x, x * F32(slope)).astype(F32)


def relu(x: np.ndarray) -> np.ndarray:
    return np.maximum(x, F32(0.0))


def avg_pool_3s2p1(x: np.ndarray) -> np.ndarray:
    """nn.AvgPool2d(3, stride=2, padding=1, count_include_pad=True): always /9.
    Condition_arch.py:10."""
    c, h, w = x.shape
    ho = (h - 1) // 2 + 1
    wo = (w - 1) // 2 + 1
    xp = np.pad(x.astype(F32), ((0, 0), (1, 1), (1, 1)))
    acc = np.zeros((c, ho, wo), dtype=F32)
    for ky in range(3):
        for kx in range(3):
            acc += xp[:, ky:ky + (ho - 1) * 2 + 1:2, kx:kx + (wo - 1) * 2 + 1:2]
    return (acc / F32(9.0)).astype(F32)


def instance_norm(x: np.ndarray, gamma: np.ndarray, beta: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """nn.InstanceNorm2d(affine=True, track_running_stats=False): per-frame,
    per-channel biased variance.  Condition_arch.py:14."""
    xm = x.astype(np.float64)
    mean = xm.mean(axis=(1, 2), keepdims=True)
    var = xm.var(axis=(1, 2), keepdims=True)
    y = (xm - mean) / np.sqrt(var + eps)
    return (y * gamma.astype(np.float64)[:, None, None] + beta.astype(np.float64)[:, None, None]).astype(F32)


def pixel_shuffle2(x: np.ndarray) -> np.ndarray:
    """nn.PixelShuffle(2): (4C,H,W) -> (C,2H,2W); out[c,2y+i,2x+j] = in[4c+2i+j,y,x]."""
    c4, h, w = x.shape
    c = c4 // 4
    return x.reshape(c, 2, 2, h, w).transpose(0, 3, 1, 4, 2).reshape(c, 2 * h, 2 * w)


def align_to(x: np.ndarray, rh: int, rw: int) -> np.ndarray:
    """HDRUNet3T1._align_to (HDRUNet3T1_arch.py:78-104): centre-crop when larger,
    replicate-pad when smaller."""
    xh, xw = x.shape[-2:]
    if xh > rh:
        top = (xh - rh) // 2
        x = x[..., top:top + rh, :]
    if xw > rw:
        left = (xw - rw) // 2
        x = x[..., :, left:left + rw]
    xh, xw = x.shape[-2:]
    ph, pw = rh - xh, rw - xw
    if ph > 0 or pw > 0:
        pt, pl = ph // 2, pw // 2
        x = np.pad(x, ((0, 0), (pt, ph - pt), (pl, pw - pl)), mode="edge")
    return x


# --------------------------------------------------------------------------
# P8  INT8 layouts: static activation fake-quantisation  (hdrtvnet_torch.py:296-410)
# --------------------------------------------------------------------------

def split_int8_state(raw: dict):
    """Eager INT8 checkpoint arrays -> (state-dict with de-quantised fp32 weights + ``<layer>.x_scale`` /
    ``<layer>.x_zero`` entries kept).  weight = weight_int8 * w_scale per output channel (:361)."""
    sd = {}
    for k, v in raw.items():
        if k.endswith(".weight_int8"):
            layer = k[: -len(".weight_int8")]
            # W8A8 layers carry ".w_scale" (:361), weight-only W8 layers ".scale" (W8Conv2d / W8Linear, :233-291)
            sc = raw[layer + ".w_scale"] if (layer + ".w_scale") in raw else raw[layer + ".scale"]
            scale = np.asarray(sc, dtype=F32).reshape((-1,) + (1,) * (v.ndim - 1))
            sd[layer + ".weight"] = (np.asarray(v, dtype=F32) * scale).astype(F32)
        elif k.endswith(".w_scale") or (k.endswith(".scale") and (k[: -len(".scale")] + ".weight_int8") in raw):
            continue
        else:
            sd[k] = np.asarray(v, dtype=F32)
    return sd


def fake_quant_input(sd: dict, layer: str, x: np.ndarray) -> np.ndarray:
    """W8A8Conv2d / W8A8Linear.forward (:350-360) on the layer input; identity when the layer has no ``x_scale``."""
    s = sd.get(layer + ".x_scale")
    if s is None:
        return x
    s = F32(np.asarray(s).reshape(-1)[0])
    z = sd.get(layer + ".x_zero")
    xf = x.astype(F32)
    if z is not None:
        z = F32(np.asarray(z).reshape(-1)[0])
        q = np.clip(np.rint((xf - z) / s), 0, 255).astype(F32)          # torch.round = round-half-even = rint
        return (q * s + z).astype(F32)
    q = np.clip(np.rint(xf / s), -128, 127).astype(F32)
    return (q * s).astype(F32)


# --------------------------------------------------------------------------
# P2  AGCM  (Condition_arch.py:8-35, 483-494, 559-585)
# --------------------------------------------------------------------------

_CLS_CONV = (0, 4, 8, 12, 16)      # nn.Sequential indices of the five 1x1 convs
_CLS_NORM = (3, 7, 11, 15, None)   # InstanceNorm after blocks 1-4, none after block 5


def classifier(sd: dict, cond_chw: np.ndarray, prefix: str = "AGCM.classifier.model.") -> np.ndarray:
    """Color_Condition: 5x [conv1x1 -> AvgPool(3,2,1) -> LeakyReLU(0.2) -> IN] ->
    Dropout(eval: identity) -> conv1x1 128->6 -> global mean.  Returns fea[6]."""
    x = cond_chw.astype(F32)
    for ci, ni in zip(_CLS_CONV, _CLS_NORM):
        x = conv2d(fake_quant_input(sd, f"{prefix}{ci}", x), sd[f"{prefix}{ci}.weight"], sd[f"{prefix}{ci}.bias"], 1, 0)
        x = avg_pool_3s2p1(x)
        x = leaky_relu(x, 0.2)
        if ni is not None:
            x = instance_norm(x, sd[f"{prefix}{ni}.weight"], sd[f"{prefix}{ni}.bias"])
    x = conv2d(fake_quant_input(sd, f"{prefix}20", x), sd[f"{prefix}20.weight"], sd[f"{prefix}20.bias"], 1, 0)
    return x.astype(np.float64).mean(axis=(1, 2)).astype(F32)


def gfm_params(sd: dict, fea: np.ndarray) -> dict:
    """Six nn.Linear(6 -> 64/64/3) heads.  Condition_arch.py:562-569."""
    out = {}
    for name in ("first", "HR", "last"):
        for kind in ("scale", "shift"):
            w = sd[f"AGCM.cond_{kind}_{name}.weight"].astype(F32)
            b = sd[f"AGCM.cond_{kind}_{name}.bias"].astype(F32)
            out[f"{kind}_{name}"] = (w @ fake_quant_input(sd, f"AGCM.cond_{kind}_{name}", fea.astype(F32)) + b).astype(F32)
    return out


def agcm(sd: dict, x_chw: np.ndarray, cond_chw: np.ndarray) -> np.ndarray:
    """ConditionNet.forward dynamic mode.  `out*scale + shift + out` per layer."""
    fea = classifier(sd, cond_chw)
    g = gfm_params(sd, fea)

    def mod(o, s, t):
        return o * s[:, None, None] + t[:, None, None] + o

    o = conv2d(fake_quant_input(sd, "AGCM.conv_first", x_chw), sd["AGCM.conv_first.weight"], sd["AGCM.conv_first.bias"], 1, 0)
    o = relu(mod(o, g["scale_first"], g["shift_first"]))
    o = conv2d(fake_quant_input(sd, "AGCM.HRconv", o), sd["AGCM.HRconv.weight"], sd["AGCM.HRconv.bias"], 1, 0)
    o = relu(mod(o, g["scale_HR"], g["shift_HR"]))
    o = conv2d(fake_quant_input(sd, "AGCM.conv_last", o), sd["AGCM.conv_last.weight"], sd["AGCM.conv_last.bias"], 1, 0)
    return mod(o, g["scale_last"], g["shift_last"]).astype(F32)


# --------------------------------------------------------------------------
# P3  LE  (HDRUNet3T1_arch.py:10-76, 152-206; arch_util.py:60-95)
# --------------------------------------------------------------------------

def _conv(sd, name, x, stride=1):
    return conv2d(fake_quant_input(sd, name, x), sd[name + ".weight"], sd[name + ".bias"], stride)


def sft_layer(sd: dict, prefix: str, fea: np.ndarray, cond: np.ndarray) -> np.ndarray:
    """SFTLayer.forward arch_util.py:68-72."""
    scale = _conv(sd, prefix + ".SFT_scale_conv1", leaky_relu(_conv(sd, prefix + ".SFT_scale_conv0", cond), 0.1))
    shift = _conv(sd, prefix + ".SFT_shift_conv1", leaky_relu(_conv(sd, prefix + ".SFT_shift_conv0", cond), 0.1))
    return (fea * (scale + F32(1.0)) + shift).astype(F32)


def resblock_sft(sd: dict, prefix: str, x: np.ndarray, cond: np.ndarray) -> np.ndarray:
    """ResBlock_with_SFT.forward arch_util.py:89-95."""
    fea = sft_layer(sd, prefix + ".sft1", x, cond)
    fea = relu(_conv(sd, prefix + ".conv1", fea))
    fea = sft_layer(sd, prefix + ".sft2", fea, cond)
    fea = _conv(sd, prefix + ".conv2", fea)
    return (x + fea).astype(F32)


def _seq(sd, prefix, x, spec):
    """spec: list of (index, stride, act_after)"""
    for idx, stride, act in spec:
        x = _conv(sd, f"{prefix}.{idx}", x, stride)
        if act:
            x = leaky_relu(x, 0.1)
    return x


def le(sd: dict, img: np.ndarray, cond_img: np.ndarray, return_intermediates: bool = False):
    """HDRUNet3T1._forward_safe_aligned with weighting_network=False, act=relu."""
    p = "LE."
    cond = _seq(sd, p + "cond_first", cond_img, [(0, 1, True), (2, 1, True), (4, 1, True)])
    cond1 = _seq(sd, p + "CondNet1", cond, [(0, 1, True), (2, 1, True), (4, 1, False)])
    cond2 = _seq(sd, p + "CondNet2", cond, [(0, 2, True), (2, 1, True), (4, 1, False)])
    cond3 = _seq(sd, p + "CondNet3", cond, [(0, 2, True), (2, 2, True), (4, 1, False)])
    cond4 = _seq(sd, p + "CondNet4", cond, [(0, 2, True), (2, 2, True), (4, 2, False)])

    fea0 = relu(_conv(sd, p + "conv_first", img))
    fea0 = sft_layer(sd, p + "SFT_layer1", fea0, cond1)
    fea0 = relu(_conv(sd, p + "HR_conv1", fea0))

    fea1 = relu(_conv(sd, p + "down_conv1", fea0, 2))
    fea1 = resblock_sft(sd, p + "recon_trunk1.0", fea1, cond2)

    fea2 = relu(_conv(sd, p + "down_conv2", fea1, 2))
    fea2 = resblock_sft(sd, p + "recon_trunk2.0", fea2, cond3)

    fea3 = relu(_conv(sd, p + "down_conv3", fea2, 2))
    out = fea3
    for i in range(4):
        out = resblock_sft(sd, f"{p}recon_trunk3.{i}", out, cond4)
    out = out + fea3

    up = relu(pixel_shuffle2(_conv(sd, p + "up_conv1.0", out)))
    up = align_to(up, *fea2.shape[-2:])
    out = resblock_sft(sd, p + "recon_trunk4.0", up + fea2, cond3)

    up = relu(pixel_shuffle2(_conv(sd, p + "up_conv2.0", out)))
    up = align_to(up, *fea1.shape[-2:])
    out = resblock_sft(sd, p + "recon_trunk5.0", up + fea1, cond2)

    up = relu(pixel_shuffle2(_conv(sd, p + "up_conv3.0", out)))
    up = align_to(up, *fea0.shape[-2:])
    out = sft_layer(sd, p + "SFT_layer2", up + fea0, cond1)

    out = relu(_conv(sd, p + "HR_conv2", out))
    out = _conv(sd, p + "conv_last", out)
    out = align_to(out, *img.shape[-2:])
    res = (img + out).astype(F32)
    if return_intermediates:
        return res, dict(cond=cond, cond1=cond1, cond2=cond2, cond3=cond3, cond4=cond4,
                         fea0=fea0, fea1=fea1, fea2=fea2, fea3=fea3)
    return res


# --------------------------------------------------------------------------
# P3'  Ensemble  (Ensemble_AGCM_LE_arch.py:889-897)
# --------------------------------------------------------------------------

def strip_module_prefix(sd: dict) -> dict:
    """hdrtvnet_torch.py:2154-2157."""
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def infer(sd: dict, x: np.ndarray, cond: np.ndarray):
    """(1,3,H,W), (1,3,h,w) -> (out (1,3,H,W), agcm_out (1,3,H,W)), fp32."""
    sd = {k: np.asarray(v, dtype=F32) for k, v in sd.items()}
    agcm_out = agcm(sd, x[0].astype(F32), cond[0].astype(F32))
    out = le(sd, agcm_out, agcm_out)
    return out[None], agcm_out[None]


# --------------------------------------------------------------------------
# P4  RGB48 pack  (gui_pipeline_worker_feeders.py:193-249)
# --------------------------------------------------------------------------

def pack_rgb48(out_1chw: np.ndarray) -> np.ndarray:
    """(1,3,H,W) any float dtype -> uint16 (H,W,3) RGB (rgb48le), FP32 math:
    clamp(0,1) -> *65535.0f (rounded) -> +0.5f (rounded) -> truncate."""
    f = out_1chw[0].astype(F32).transpose(1, 2, 0)
    f = np.clip(f, F32(0.0), F32(1.0))
    f = f * F32(65535.0)
    f = f + F32(0.5)
    return f.astype(np.uint16)


_PQ_M1 = 2610.0 / 16384.0
_PQ_M2 = 2523.0 / 32.0
_PQ_C1 = 3424.0 / 4096.0
_PQ_C2 = 2413.0 / 128.0
_PQ_C3 = 2392.0 / 128.0


def pq_oetf_absolute(luminance: np.ndarray) -> np.ndarray:
    """gui_objective_metrics.py:486-491 (_pq_oetf_absolute)."""
    y = np.clip(luminance.astype(F32, copy=False) / 10000.0, 0.0, 1.0)
    y_m1 = np.power(y, _PQ_M1).astype(F32, copy=False)
    num = _PQ_C1 + (_PQ_C2 * y_m1)
    den = 1.0 + (_PQ_C3 * y_m1)
    return np.power(num / np.maximum(den, 1e-12), _PQ_M2).astype(F32, copy=False)


def pack_rgb48_pq(linear_1chw: np.ndarray, peak_nits: float = 1000.0) -> np.ndarray:
    """Optional transfer: treat the tensor as linear light in [0,1] of `peak_nits`
    and PQ-encode it.  Follows _linear_bgr_to_bt2100_pq_bgr_u16
    (gui_objective_metrics.py:531-539) but keeps RGB channel order."""
    rgb = np.clip(linear_1chw[0].astype(F32).transpose(1, 2, 0), 0.0, 1.0) * float(peak_nits)
    pq = pq_oetf_absolute(rgb.astype(F32))
    return np.clip((pq * 65535.0) + 0.5, 0.0, 65535.0).astype(np.uint16)


# --------------------------------------------------------------------------
# P5  BGR24 postprocess  (hdrtvnet_torch.py:2352-2368)
# --------------------------------------------------------------------------

def postprocess_bgr24(out_1chw: np.ndarray) -> np.ndarray:
    """clamp_(0,1).mul_(255).add_(0.5) IN THE TENSOR'S OWN DTYPE (each op rounds
    to that dtype), truncate to u8, RGB->BGR, HWC."""
    dt = out_1chw.dtype if out_1chw.dtype in (np.float16, np.float32) else F32
    t = out_1chw[0].astype(dt)
    t = np.clip(t, dt.type(0.0), dt.type(1.0))
    t = (t * dt.type(255.0)).astype(dt)
    t = (t + dt.type(0.5)).astype(dt)
    u = t.astype(F32).astype(np.uint8)
    return np.ascontiguousarray(u[::-1].transpose(1, 2, 0))


def process(sd: dict, frame_bgr: np.ndarray):
    """HDRTVNetTorch.process in fp32: BGR u8 -> BGR u8."""
    x, cond = preprocess(frame_bgr, np.float32)
    out, _ = infer(sd, x, cond)
    return postprocess_bgr24(out)


# --------------------------------------------------------------------------
# Weights helper: seeded random state-dict with the reference's key set/shapes
# (values are the oracle's own; the reference-initialised set lives in
# tests/golden/).
# --------------------------------------------------------------------------

def state_dict_spec():
    spec = {}
    cls = "AGCM.classifier.model."
    chans = [3, 16, 32, 64, 128, 128]
    for i, ci in enumerate(_CLS_CONV):
        spec[f"{cls}{ci}.weight"] = (chans[i + 1], chans[i], 1, 1)
        spec[f"{cls}{ci}.bias"] = (chans[i + 1],)
        if _CLS_NORM[i] is not None:
            spec[f"{cls}{_CLS_NORM[i]}.weight"] = (chans[i + 1],)
            spec[f"{cls}{_CLS_NORM[i]}.bias"] = (chans[i + 1],)
    spec[f"{cls}20.weight"] = (6, 128, 1, 1)
    spec[f"{cls}20.bias"] = (6,)
    for name, n in (("first", 64), ("HR", 64), ("last", 3)):
        for kind in ("scale", "shift"):
            spec[f"AGCM.cond_{kind}_{name}.weight"] = (n, 6)
            spec[f"AGCM.cond_{kind}_{name}.bias"] = (n,)
    spec["AGCM.conv_first.weight"] = (64, 3, 1, 1)
    spec["AGCM.HRconv.weight"] = (64, 64, 1, 1)
    spec["AGCM.conv_last.weight"] = (3, 64, 1, 1)

    def conv(name, o, c, k):
        spec[name + ".weight"] = (o, c, k, k)
        spec[name + ".bias"] = (o,)

    for n, o in (("AGCM.conv_first", 64), ("AGCM.HRconv", 64), ("AGCM.conv_last", 3)):
        spec[n + ".bias"] = (o,)

    def sft(prefix):
        conv(prefix + ".SFT_scale_conv0", 16, 16, 1)
        conv(prefix + ".SFT_scale_conv1", 32, 16, 1)
        conv(prefix + ".SFT_shift_conv0", 16, 16, 1)
        conv(prefix + ".SFT_shift_conv1", 32, 16, 1)

    conv("LE.conv_first", 32, 3, 3)
    sft("LE.SFT_layer1")
    conv("LE.HR_conv1", 32, 32, 3)
    for i in (1, 2, 3):
        conv(f"LE.down_conv{i}", 32, 32, 3)
        conv(f"LE.up_conv{i}.0", 128, 32, 3)
    for t, n in ((1, 1), (2, 1), (3, 4), (4, 1), (5, 1)):
        for j in range(n):
            pre = f"LE.recon_trunk{t}.{j}"
            conv(pre + ".conv1", 32, 32, 3)
            conv(pre + ".conv2", 32, 32, 3)
            sft(pre + ".sft1")
            sft(pre + ".sft2")
    sft("LE.SFT_layer2")
    conv("LE.HR_conv2", 32, 32, 3)
    conv("LE.conv_last", 3, 32, 3)
    conv("LE.cond_first.0", 64, 3, 3)
    conv("LE.cond_first.2", 64, 64, 1)
    conv("LE.cond_first.4", 64, 64, 1)
    conv("LE.CondNet1.0", 64, 64, 1)
    conv("LE.CondNet1.2", 64, 64, 1)
    conv("LE.CondNet1.4", 16, 64, 1)
    conv("LE.CondNet2.0", 64, 64, 3)
    conv("LE.CondNet2.2", 64, 64, 1)
    conv("LE.CondNet2.4", 16, 64, 1)
    conv("LE.CondNet3.0", 64, 64, 3)
    conv("LE.CondNet3.2", 64, 64, 3)
    conv("LE.CondNet3.4", 16, 64, 1)
    conv("LE.CondNet4.0", 64, 64, 3)
    conv("LE.CondNet4.2", 64, 64, 3)
    conv("LE.CondNet4.4", 16, 64, 3)
    return spec


def random_state_dict(seed: int = 0) -> dict:
    """Seeded weights with the reference key set.  Kaiming-uniform-like fan-in
    scaling so activations stay O(1); resblock convs scaled by 0.1 as
    arch_util.py:87 does; InstanceNorm affine = (1 + small, small)."""
    rng = np.random.default_rng(seed)
    sd = {}
    for k, shp in state_dict_spec().items():
        if k.endswith(".weight") and len(shp) == 1:      # InstanceNorm gamma
            sd[k] = (1.0 + 0.1 * rng.standard_normal(shp)).astype(F32)
        elif k.endswith(".bias"):
            sd[k] = (0.05 * rng.standard_normal(shp)).astype(F32)
        else:
            fan_in = int(np.prod(shp[1:]))
            bound = np.sqrt(3.0 / fan_in)
            w = rng.uniform(-bound, bound, size=shp).astype(F32)
            if "recon_trunk" in k and (".conv1." in k or ".conv2." in k):
                w *= F32(0.3)
            sd[k] = w
    return sd


# --------------------------------------------------------------------------
# HG stage: Hallucination_Generator + HG_Composite
# (src/models/hdrtvnet_modules/Hallucination_arch.py:53-137,
#  src/models/hdrtvnet_modules/HG_Composite_arch.py:77-107)
# --------------------------------------------------------------------------

HG_BN_BLOCKS = ("conv1", "conv2", "conv3_1", "conv3_2", "conv4_1", "conv4_2", "conv5_1", "conv5_2", "conv_code1",
                "conv_code2")


def max_pool2(x: np.ndarray) -> np.ndarray:
    """nn.MaxPool2d(2) on (C,H,W) with even H, W (Hallucination_arch.py:57)."""
    c, h, w = x.shape
    return x[:, :h // 2 * 2, :w // 2 * 2].reshape(c, h // 2, 2, w // 2, 2).max(axis=(2, 4))


def batch_norm_eval(x: np.ndarray, sd: dict, prefix: str, eps: float = 1e-5) -> np.ndarray:
    """nn.BatchNorm2d in eval mode: (x - running_mean) / sqrt(running_var + eps) * weight + bias
    (conv_block, Hallucination_arch.py:24-29)."""
    mean = sd[prefix + ".running_mean"].astype(F32)[:, None, None]
    var = sd[prefix + ".running_var"].astype(F32)[:, None, None]
    g = sd[prefix + ".weight"].astype(F32)[:, None, None]
    b = sd[prefix + ".bias"].astype(F32)[:, None, None]
    return ((x - mean) / np.sqrt(var + F32(eps)) * g + b).astype(F32)


def hg_fold_bn(sd: dict, eps: float = 1e-5) -> dict:
    """Eval-time BatchNorm folded into the preceding conv, the arithmetic of
    Hallucination_Generator_FusedBN._fold_bn_into_conv (Hallucination_arch.py:240-275):
    scale = bn_w * rsqrt(var + eps); w' = w * scale; b' = (b - mean) * scale + bn_b."""
    out = {k: np.asarray(v) for k, v in sd.items()}
    for blk in HG_BN_BLOCKS:
        if f"{blk}.1.running_var" not in out:
            continue
        scale = out[f"{blk}.1.weight"].astype(F32) / np.sqrt(out[f"{blk}.1.running_var"].astype(F32) + F32(eps))
        out[f"{blk}.0.weight"] = (out[f"{blk}.0.weight"].astype(F32) * scale[:, None, None, None]).astype(F32)
        out[f"{blk}.0.bias"] = ((out[f"{blk}.0.bias"].astype(F32) - out[f"{blk}.1.running_mean"].astype(F32)) * scale
                                + out[f"{blk}.1.bias"].astype(F32)).astype(F32)
        for s in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
            out.pop(f"{blk}.1.{s}", None)
    return out


def hg_mask(img_chw: np.ndarray, r: float = 0.75, thresh: float = 0.1) -> np.ndarray:
    """HG_Composite._make_mask (HG_Composite_arch.py:77-84): 1 where (max_c(img) - r) / (1 - r), clamped to
    [0, 1], exceeds thresh.  fp32 arithmetic with the python-double constants rounded to fp32 as torch does."""
    m = img_chw.astype(F32).max(axis=0, keepdims=True)
    m = (m - F32(r)) / F32(1.0 - r)
    m = np.clip(m, F32(0.0), F32(1.0))
    return (m > F32(thresh)).astype(F32)


def hg_unet(sd: dict, img: np.ndarray, return_intermediates: bool = False):
    """Hallucination_Generator.forward without the final mask blend (Hallucination_arch.py:101-135):
    (3,H,W) with H, W multiples of 32 -> conv_last(cat(conv10_out, img)) (3,H,W)."""
    def block(name, x):
        y = conv2d(x, sd[name + ".0.weight"], sd[name + ".0.bias"])
        if name + ".1.running_var" in sd:
            y = batch_norm_eval(y, sd, name + ".1")
        return relu(y)

    def up(name, x):
        return relu(pixel_shuffle2(conv2d(x, sd[name + ".0.weight"], sd[name + ".0.bias"])))

    def c1x1(name, x):
        return conv2d(x, sd[name + ".weight"], sd[name + ".bias"])

    t = {}
    t["conv1"] = block("conv1", img)
    t["conv2"] = block("conv2", max_pool2(t["conv1"]))
    t["conv3"] = block("conv3_2", max_pool2(block("conv3_1", t["conv2"])))
    t["conv4"] = block("conv4_2", max_pool2(block("conv4_1", t["conv3"])))
    t["conv5"] = block("conv5_2", max_pool2(block("conv5_1", t["conv4"])))
    t["code"] = block("conv_code2", max_pool2(block("conv_code1", t["conv5"])))
    t["conv6"] = c1x1("conv6", np.concatenate((up("Up_conv1", t["code"]), t["conv5"]), axis=0))
    t["conv7"] = c1x1("conv7", np.concatenate((up("Up_conv2", t["conv6"]), t["conv4"]), axis=0))
    t["conv8"] = c1x1("conv8", np.concatenate((up("Up_conv3", t["conv7"]), t["conv3"]), axis=0))
    t["conv9"] = c1x1("conv9", np.concatenate((up("Up_conv4", t["conv8"]), t["conv2"]), axis=0))
    t["conv10"] = c1x1("conv10", np.concatenate((up("Up_conv5", t["conv9"]), t["conv1"]), axis=0))
    out = c1x1("conv_last", np.concatenate((t["conv10"], img.astype(F32)), axis=0))
    return (out, t) if return_intermediates else out


def hg_stage(hg_sd: dict, base_out_1chw: np.ndarray, mask_r: float = 0.75):
    """HG_Composite.forward after the base model (HG_Composite_arch.py:88-107): mask from the base output,
    reflect-pad right / bottom to the next multiple of 32, HG, crop.  (1,3,H,W) fp32 -> (1,3,H,W) fp32."""
    sd = {k: np.asarray(v) for k, v in hg_sd.items()}
    base = base_out_1chw[0].astype(F32)
    _, h, w = base.shape
    mask = hg_mask(base, r=mask_r)
    ph, pw = (32 - h % 32) % 32, (32 - w % 32) % 32
    img = np.pad(base, ((0, 0), (0, ph), (0, pw)), mode="reflect") if (ph or pw) else base
    mpad = np.pad(mask, ((0, 0), (0, ph), (0, pw)), mode="reflect") if (ph or pw) else mask
    out = mpad * hg_unet(sd, img) + img            # Hallucination_arch.py:136
    return out[None, :, :h, :w].astype(F32)


# --------------------------------------------------------------------------
# Letterbox (second half): src/gui_scaling.py:228-244 `_letterbox_bgr`
#   scale = min(out_w / w, out_h / h); new = round(size * scale); cv2.resize with INTER_AREA when shrinking, INTER_CUBIC
#   when enlarging; centred on a black canvas.
# cv2.resize is third-party (opencv-python 4.13 here); its published algorithm (modules/imgproc/src/resize.cpp) restated:
#   INTER_AREA, integer factors : block sums; 2x2 -> (s + 2) >> 2, else saturate(rint(float(s) * float(1 / area)))
#   INTER_AREA, general         : computeResizeAreaTab weights (fp32), horizontal then vertical accumulation in fp32, in
#                                 table order, products rounded before they are added (no FMA), saturate(rint(.))
#   INTER_CUBIC (8-bit)         : a = -0.75 cubic weights in fp32 -> short fixed point (x 2048, round-half-even);
#                                 horizontal pass exact in int32 on replicate-clamped taps; vertical pass in fp32
#                                 (weights x 2^-22): S3*b3, then S2*b2 + ., S1*b1 + ., S0*b0 + . (no FMA), rint, saturate
# Pinned in tests/test_oracle_golden.py against cv2 itself: INTER_AREA bit-exact with and without IPP; INTER_CUBIC
# bit-exact against OpenCV's own code path (cv2.ipp.setUseIPP(False)) and within 1 code of the closed-source Intel IPP
# routine the pip wheel dispatches to by default.
# --------------------------------------------------------------------------

def letterbox_geometry(h: int, w: int, out_h: int, out_w: int):
    """(new_h, new_w, y0, x0, shrink) of gui_scaling.py:234-243; python round() = round-half-even on the double."""
    scale = min(out_w / max(w, 1), out_h / max(h, 1))
    new_w = max(1, int(round(w * scale)))
    new_h = max(1, int(round(h * scale)))
    return new_h, new_w, (out_h - new_h) // 2, (out_w - new_w) // 2, scale < 1.0


def _area_table(ssize: int, dsize: int):
    """computeResizeAreaTab: list of (dst index, src index, fp32 weight)."""
    import math
    scale = 1.0 / (np.float64(dsize) / np.float64(ssize))
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, F32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, F32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, F32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def resize_area(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    h, w, cn = img.shape
    if h % new_h == 0 and w % new_w == 0:                                   # resizeAreaFast_
        sy, sx = h // new_h, w // new_w
        s = img.astype(np.int64).reshape(new_h, sy, new_w, sx, cn).sum((1, 3))
        if sx == 2 and sy == 2:
            return ((s + 2) >> 2).astype(np.uint8)
        return np.clip(np.rint(s.astype(F32) * F32(1.0 / (sx * sy))), 0, 255).astype(np.uint8)
    xt, yt = _area_table(w, new_w), _area_table(h, new_h)
    src = img.astype(F32)
    buf = np.zeros((h, new_w, cn), F32)
    for dx, sx, a in xt:
        buf[:, dx, :] = buf[:, dx, :] + (src[:, sx, :] * a).astype(F32)
    out = np.zeros((new_h, new_w, cn), np.uint8)
    prev, acc = -1, None
    for dy, sy, b in yt:
        if dy != prev:
            if prev >= 0:
                out[prev] = np.clip(np.rint(acc), 0, 255)
            acc, prev = (buf[sy] * b).astype(F32), dy
        else:
            acc = acc + (buf[sy] * b).astype(F32)
    out[prev] = np.clip(np.rint(acc), 0, 255)
    return out


def _cubic_table(n_src: int, n_dst: int):
    """Tap indices (replicate-clamped) and short fixed-point weights of the 8-bit INTER_CUBIC path."""
    scale = 1.0 / (np.float64(n_dst) / np.float64(n_src))
    f = ((np.arange(n_dst, dtype=np.float64) + 0.5) * scale - 0.5).astype(F32)
    s = np.floor(f).astype(np.int32)
    x = (f - s.astype(F32)).astype(F32)
    a = F32(-0.75)
    one, two, three = F32(1), F32(2), F32(3)
    c0 = ((a * (x + one) - F32(5) * a) * (x + one) + F32(8) * a) * (x + one) - F32(4) * a
    c1 = ((a + two) * x - (a + three)) * x * x + one
    y = one - x
    c2 = ((a + two) * y - (a + three)) * y * y + one
    c3 = one - c0 - c1 - c2
    ic = np.rint(np.stack([c0, c1, c2, c3], -1).astype(F32) * F32(2048)).astype(np.int32)
    idx = np.clip(s[:, None] + np.arange(-1, 3)[None, :], 0, n_src - 1)
    return idx, ic


def resize_cubic(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    xi, xa = _cubic_table(img.shape[1], new_w)
    yi, ya = _cubic_table(img.shape[0], new_h)
    hor = (img.astype(np.int64)[:, xi, :] * xa[None, :, :, None]).sum(2)              # exact integers < 2^24
    b = (ya.astype(F32) * F32(1.0 / (2048 * 2048))).astype(F32)
    rows = hor[yi, :, :]                                                              # (new_h, 4, new_w, cn)
    fl = rows.astype(F32)
    acc = (fl[:, 3] * b[:, 3, None, None]).astype(F32)
    for k in (2, 1, 0):
        acc = ((fl[:, k] * b[:, k, None, None]).astype(F32) + acc).astype(F32)
    out = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    # scalar tail of OpenCV's vector loop (8 int16 lanes per step over width * channels): fixed point
    flat = new_w * img.shape[2]
    tail = flat - flat // 8 * 8
    if tail:
        fx = (rows * ya[:, :, None, None]).sum(1)
        fixed = np.clip((fx + (1 << 21)) >> 22, 0, 255).astype(np.uint8).reshape(new_h, flat)
        out = out.reshape(new_h, flat)
        out[:, flat - tail:] = fixed[:, flat - tail:]
        out = out.reshape(new_h, new_w, img.shape[2])
    return out


def letterbox_bgr(frame: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """gui_scaling.py:228-244."""
    h, w = frame.shape[:2]
    if w == out_w and h == out_h:
        return frame
    new_h, new_w, y0, x0, shrink = letterbox_geometry(h, w, out_h, out_w)
    if new_h == h and new_w == w:
        resized = frame                                                               # cv2.resize to the same size copies
    else:
        resized = resize_area(frame, new_w, new_h) if shrink else resize_cubic(frame, new_w, new_h)
    canvas = np.zeros((out_h, out_w, 3), dtype=frame.dtype)
    canvas[y0:y0 + new_h, x0:x0 + new_w] = resized
    return canvas
