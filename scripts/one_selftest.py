#!/usr/bin/env python
"""Run one conv self-test (for compute-sanitizer): python scripts/one_selftest.py kind cin cout H W flags"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import hdr_realtime_video_pipeline_b200 as hb  # noqa: E402

args = [int(a) for a in sys.argv[1:7]] if len(sys.argv) >= 7 else [1, 16, 16, 8, 128, 0]
net = hb.HDRTVNetB200(os.path.join(REPO, "tests/golden/weights_hr.npz"), precision="fp16", warmup_passes=0, use_hg=False, debug_library=True)
print(net.conv_selftest(*args))
