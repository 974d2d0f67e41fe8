"""GPU letterbox (hdrtv_letterbox_bgr; reference: src/gui_scaling.py:228-244 `_letterbox_bgr`) against the oracle - which
tests/test_oracle_golden.py pins against cv2.resize itself - and, where cv2 is importable, against cv2 directly.
uint8 in, uint8 out: the bar is bit-exact."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

import hdr_realtime_video_pipeline_b200 as hb  # noqa: E402
from oracle import hdrtvnet_oracle as O  # noqa: E402

CASES = [  # (h, w, out_w, out_h)
    (2160, 3840, 1920, 1080),    # INTER_AREA 2x2 fast path (4K source, 1080p processing)
    (1080, 1920, 640, 360),      # INTER_AREA 3x3 fast path
    (1080, 1920, 1280, 720),     # INTER_AREA general (1.5x)
    (1080, 1440, 1280, 720),     # 4:3 source: pillarbox, general area
    (800, 1920, 1280, 720),      # 2.4:1 source: letterbox bars
    (720, 1280, 1920, 1080),     # INTER_CUBIC 1.5x
    (540, 960, 1920, 1080),      # INTER_CUBIC 2x
    (480, 640, 1920, 1080),      # cubic + pillarbox
    (123, 217, 333, 251),        # odd sizes (vector-loop tail: 3 * new_w not a multiple of 8)
    (1080, 1920, 1920, 1080),    # same size: returned as is
    (1080, 1920, 2000, 1200),    # canvas larger with the same scale-1 frame: copy + bars
]


@pytest.fixture(scope="module")
def net():
    n = hb.HDRTVNetB200(os.path.join(GOLDEN, "weights_hr.npz"), device="cuda", precision="fp16", warmup_passes=0, use_hg=False)
    yield n
    n.close()


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_to_%dx%d" % c)
def test_letterbox_is_bit_exact(net, case):
    h, w, out_w, out_h = case
    frame = np.random.default_rng(h + 3 * w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    want = O.letterbox_bgr(frame, out_w, out_h)
    got = net.letterbox_bgr(frame, out_w, out_h)
    assert got.dtype == np.uint8 and got.shape == (out_h, out_w, 3)
    assert np.array_equal(got, want)
    dev = net.letterbox_bgr(torch.from_numpy(frame).cuda(), out_w, out_h)        # device in -> device out
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), want)
    try:
        import cv2
    except Exception:
        return
    ipp = cv2.ipp.useIPP()
    try:
        cv2.ipp.setUseIPP(False)
        new_h, new_w, y0, x0, shrink = O.letterbox_geometry(h, w, out_h, out_w)
        if (new_h, new_w) != (h, w):
            ref = cv2.resize(frame, (new_w, new_h), interpolation=cv2.INTER_AREA if shrink else cv2.INTER_CUBIC)
            assert np.array_equal(got[y0:y0 + new_h, x0:x0 + new_w], ref)
    finally:
        cv2.ipp.setUseIPP(ipp)


def test_letterboxed_frame_feeds_the_one_call_path(net):
    """4K source on a 1080p processing size: letterbox on the device, then process_rgb48 on the device tensor = the
    reference's host sequence (_letterbox_bgr -> preprocess -> infer -> RGB48 pack)."""
    src = hb.synth_frame(1, 432, 768, "ramps")
    small = net.letterbox_bgr(torch.from_numpy(src).cuda(), 384, 216)
    host = O.letterbox_bgr(src, 384, 216)
    assert np.array_equal(small.cpu().numpy(), host)
    a = net.process_rgb48(small, serial=True)
    a.wait_ready()
    got = a.numpy().copy()
    a.release()
    out, _ = net.infer(net.preprocess(host))
    fr = hb.tensor_to_rgb48_bytes(out, {})
    want = fr.numpy().copy()
    fr.release()
    assert np.array_equal(got, want)
