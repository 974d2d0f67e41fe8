"""Deterministic synthetic SDR frames (decode stand-in) for tests and benchmarks.

Four content classes cycled by frame index:
  A  i.i.d. uniform noise, full code range
  B  smooth ramps + a diagonal term that moves with the frame index (video-like)
  C  all-zero frame (the reference's own micro-benchmark input,
     src/build_tensorrt_engines.py:541)
  D  all-255 with 5 % salt noise (drives outputs to the top of the range)
Frames are uint8 HxWx3 BGR, C-contiguous, as cv2 decode would deliver them
(src/video_source.py).
"""
from __future__ import annotations

import numpy as np

CLASSES = ("noise", "ramps", "black", "white_salt")


def synth_frame(frame_idx: int, height: int, width: int, cls: str | None = None) -> np.ndarray:
    rng = np.random.default_rng(1234 + int(frame_idx))
    if cls is None:
        cls = CLASSES[int(frame_idx) % len(CLASSES)]
    if cls == "noise":
        return rng.integers(0, 256, size=(height, width, 3), dtype=np.uint8)
    if cls == "ramps":
        x = np.arange(width, dtype=np.int64)[None, :]
        y = np.arange(height, dtype=np.int64)[:, None]
        b = (x * 255) // max(width - 1, 1) + 0 * y
        g = (y * 255) // max(height - 1, 1) + 0 * x
        r = (((x + y + 8 * int(frame_idx)) * 255) // max(width + height - 2, 1)) % 256
        return np.ascontiguousarray(np.stack([b, g, r], axis=-1).astype(np.uint8))
    if cls == "black":
        return np.zeros((height, width, 3), dtype=np.uint8)
    if cls == "white_salt":
        frame = np.full((height, width, 3), 255, dtype=np.uint8)
        mask = rng.random((height, width)) < 0.05
        salt = rng.integers(0, 256, size=(int(mask.sum()), 3), dtype=np.uint8)
        frame[mask] = salt
        return frame
    raise ValueError(f"unknown synthetic frame class {cls!r}")


def synth_clip(n_frames: int, height: int, width: int, first_frame: int = 0):
    for i in range(first_frame, first_frame + n_frames):
        yield i, synth_frame(i, height, width)


# ---------------------------------------------------------------------------------------------------------------------
# Seeded stand-in for HG.pt.  The reference tree does not ship the highlight-generation checkpoint
# (.MISSING_LARGE_BLOBS), so tests and benchmarks use random-init weights with the key set / shapes of
# Hallucination_Generator (Hallucination_arch.py:53-98, nf = 64): Kaiming-normal conv weights (fan-in, as
# weights_init_kaiming :13-21), BatchNorm affine N(1, 0.02) / 0 with NON-trivial running statistics (so that the
# eval-mode fold is exercised), and a down-scaled output head so that the hallucinated residual stays a small
# correction like a trained model's.  numpy's PCG64 stream is platform-independent: the build container (fixtures) and
# the GPU box regenerate identical tensors from the seed.
# ---------------------------------------------------------------------------------------------------------------------
HG_BN_BLOCKS = ("conv1", "conv2", "conv3_1", "conv3_2", "conv4_1", "conv4_2", "conv5_1", "conv5_2", "conv_code1",
                "conv_code2")


def hg_state_dict_spec(nf: int = 64) -> dict:
    spec = {}

    def conv(name, o, c, k):
        spec[name + ".weight"] = (o, c, k, k)
        spec[name + ".bias"] = (o,)

    chans = {"conv1": (3, nf), "conv2": (nf, 2 * nf), "conv3_1": (2 * nf, 4 * nf), "conv3_2": (4 * nf, 4 * nf),
             "conv4_1": (4 * nf, 8 * nf), "conv4_2": (8 * nf, 8 * nf), "conv5_1": (8 * nf, 8 * nf),
             "conv5_2": (8 * nf, 8 * nf), "conv_code1": (8 * nf, 8 * nf), "conv_code2": (8 * nf, 8 * nf)}
    for blk in HG_BN_BLOCKS:
        ci, co = chans[blk]
        conv(blk + ".0", co, ci, 3)
        for s in ("weight", "bias", "running_mean", "running_var"):
            spec[f"{blk}.1.{s}"] = (co,)
        spec[f"{blk}.1.num_batches_tracked"] = ()
    for name, ci, co in (("Up_conv1", 8 * nf, 8 * nf), ("Up_conv2", 8 * nf, 8 * nf), ("Up_conv3", 4 * nf, 4 * nf),
                         ("Up_conv4", 2 * nf, 2 * nf), ("Up_conv5", nf, nf)):
        conv(name + ".0", 4 * co, ci, 3)
    conv("conv6", 8 * nf, 16 * nf, 1)
    conv("conv7", 4 * nf, 16 * nf, 1)
    conv("conv8", 2 * nf, 8 * nf, 1)
    conv("conv9", nf, 4 * nf, 1)
    conv("conv10", 3, 2 * nf, 1)
    conv("conv_last", 3, 6, 1)
    return spec


def hg_random_state_dict(seed: int = 0, head_scale: float = 0.01) -> dict:
    rng = np.random.default_rng(77000 + int(seed))
    sd = {}
    for k, shp in hg_state_dict_spec().items():
        if k.endswith("num_batches_tracked"):
            sd[k] = np.asarray(100, dtype=np.int64)
        elif k.endswith(".1.weight"):
            sd[k] = (1.0 + 0.02 * rng.standard_normal(shp)).astype(np.float32)
        elif k.endswith(".1.bias"):
            sd[k] = (0.02 * rng.standard_normal(shp)).astype(np.float32)
        elif k.endswith("running_mean"):
            sd[k] = (0.1 * rng.standard_normal(shp)).astype(np.float32)
        elif k.endswith("running_var"):
            sd[k] = rng.uniform(0.6, 1.6, size=shp).astype(np.float32)
        elif k.endswith(".bias"):
            sd[k] = (0.05 * rng.standard_normal(shp)).astype(np.float32)
        else:
            fan_in = int(np.prod(shp[1:]))
            w = (np.sqrt(2.0 / fan_in) * rng.standard_normal(shp)).astype(np.float32)
            if k.startswith("conv_last") or k.startswith("conv10"):
                w *= np.float32(np.sqrt(head_scale))
            sd[k] = w
    return sd
