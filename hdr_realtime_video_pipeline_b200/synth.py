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
