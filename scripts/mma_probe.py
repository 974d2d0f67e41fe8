#!/usr/bin/env python
"""tcgen05 issue-rate probe: cycles per M=128 x N x K=16 kind::f16 MMA for several N / operand layouts /
accumulator counts / CTA counts (design evidence for DESIGN.md)."""
import os
import sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import hdr_realtime_video_pipeline_b200 as hb  # noqa: E402
net = hb.HDRTVNetB200(os.path.join(REPO, "tests/golden/weights_hr.npz"), precision="fp16", warmup_passes=0, use_hg=False, debug_library=True)
names = {0: "none/plane-pitch", 1: "none/dense", 2: "swizzle128"}
for blocks in (1, 148, 296):
    for layout in (0, 2):
        for nacc in (1, 2, 4):
            row = []
            for n in (16, 32, 64, 128, 256):
                if nacc * n > 512:
                    continue
                vary = 1 if layout != 2 else 0
                net.mma_probe(n, layout, vary, 200, blocks, nacc)
                row.append(f"N={n}:{net.mma_probe(n, layout, vary, 4000, blocks, nacc):6.1f}")
            print(f"blocks={blocks:3d} layout={names[layout]:18s} nacc={nacc} " + "  ".join(row), flush=True)
