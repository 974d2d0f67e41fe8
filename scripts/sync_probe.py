#!/usr/bin/env python
"""Latency of the per-row synchronisation primitives (probe kind 8): cycles per op, one warp, dependent chain."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import hdr_realtime_video_pipeline_b200 as hb  # noqa: E402

net = hb.HDRTVNetB200(os.path.join(REPO, "tests/golden/weights_hr.npz"), precision="fp16", warmup_passes=0, use_hg=False, debug_library=True)
ops = {0: "empty loop", 1: "try_wait (complete barrier, whole warp)", 2: "test_wait (complete barrier)", 3: "tcgen05.fence::after_thread_sync",
       4: "elect.sync + syncwarp", 5: "mbar_wait wrapper", 6: "lane-0 clock64 stamp to global", 7: "syncwarp + lane-0 mbarrier.arrive",
       8: "elected tcgen05.commit (nothing pending)", 9: "fence.proxy.async.shared::cta", 10: "tcgen05.fence::before_thread_sync",
       11: "lane-0 try_wait + syncwarp"}
for op, name in ops.items():
    c = net.probe(8, iters=512, groups=op, trace=(op == 6))
    c = c[0] if isinstance(c, tuple) else c
    print(f"{c:8.1f} cycles  {name}")
print("lean batches of 5 MMAs (M=128, K=16), cycles per MMA; A operand pattern: distinct tiles / same tile / dx-shifted")
for n in (16, 32, 48, 64, 96, 128, 192, 256):
    row = []
    for pat in (0, 1, 2):
        row.append(net.probe(7, n=n, iters=2000, nmma=pat, groups=0) / 5)
    print(f"  N={n:3d}: " + "  ".join(f"{v:6.1f}" for v in row))
print("same, but consecutive MMAs alternate between two accumulators (no accumulate-after-accumulate dependency)")
for n in (16, 32, 64, 96, 128):
    print(f"  N={n:3d}: {net.probe(7, n=n, iters=2000, nmma=0, groups=64) / 5:6.1f}")
print("unrolled batches of 12 MMAs, cycles per MMA: same A,B / distinct A / distinct B / distinct A and B / distinct A,B + alternating accumulators")
for n in (16, 32, 48, 64, 96, 128):
    row = [net.probe(9, n=n, iters=500, nmma=f) / 12 for f in (0, 1, 2, 3, 7)]
    print(f"  N={n:3d}: " + "  ".join(f"{v:6.1f}" for v in row))
