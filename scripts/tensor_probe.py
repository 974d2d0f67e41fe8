#!/usr/bin/env python
"""Tensor-path micro-probes (csrc/probes.cuh) — cycles per iteration on the B200.
    python scripts/tensor_probe.py"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import hdr_realtime_video_pipeline_b200 as hb  # noqa: E402

net = hb.HDRTVNetB200(os.path.join(REPO, "tests/golden/weights_hr.npz"), precision="fp16", warmup_passes=0, use_hg=False, debug_library=True)
for blocks in (1, 148):
    for kind, label in ((0, "mma SS M=128"), (1, "mma TS (A in TMEM) M=128"), (2, "mma SS M=64")):
        row = []
        for n in (16, 32, 64, 128, 256):
            row.append(f"N={n}: {net.probe(kind, n=n, iters=4000, blocks=blocks):6.1f}")
        print(f"blocks={blocks:3d} {label:26s} " + "  ".join(row), flush=True)
for blocks in (148, 296):
    row = [f"N={n}: {net.probe(0, n=n, iters=4000, blocks=blocks):6.1f}" for n in (32, 64, 128)]
    print(f"blocks={blocks:3d} (CTAs/SM = {blocks // 148}) mma SS M=128 per-CTA cycles/MMA  " + "  ".join(row), flush=True)
for kind, who in ((5, "one thread"), (6, "uniform warp + elect")):
    for flags, label in ((0, "no commit"), (1, "commit"), (17, "commit + try_wait"), (16, "try_wait only")):
        row = [f"nmma={m}: {net.probe(kind, iters=2000, nmma=m, groups=flags):7.1f}" for m in (1, 2, 5, 10, 19)]
        print(f"free-running N=64 per iteration [{who:20s} {label:18s}] " + "  ".join(row), flush=True)
for blocks in (1,):
    for shape, label in ((64, "x32"), (16, "x16")):
        row = []
        for nw in (1, 4, 8, 16):
            cyc = net.probe(3, n=shape, iters=2000, blocks=blocks, nwarps=nw)
            row.append(f"warps={nw}: {cyc:7.1f} cyc/iter = {nw * 32 * 64 * 4 / cyc:6.1f} B/cyc")
        print(f"blocks={blocks:3d} tcgen05.ld {label:4s} 64 cols/warp/iter  " + "  ".join(row), flush=True)
for flags in (0, 1, 3, 7, 15, 31, 63, 16, 8, 32):
    print(f"lean 5-MMA batches (N=64) flags={flags:2d} [1 commit 2 try_wait 4 fence 8 altD 16 acc0 32 lds-desc]: "
          f"{net.probe(7, iters=2000, groups=flags):7.1f} cycles/iteration", flush=True)
for mode, mlabel in ((0, "spin"), (1, "hinted try_wait"), (2, "nanosleep(64)")):
    for groups in (1, 2, 4):
        row = []
        for nmma in (1, 5, 10):
            row.append(f"nmma={nmma}: {net.probe(4, n=64, iters=1000, nmma=nmma, groups=groups, nwarps=mode):7.1f}")
        row.append(f"no-epilogue nmma=5: {net.probe(4, n=0, iters=1000, nmma=5, groups=groups, nwarps=mode):7.1f}")
        print(f"chain round trip [{mlabel:16s}] groups={groups} (cycles per iteration of all groups)  " + "  ".join(row), flush=True)
for mode in (0, 1):
    for groups, n in ((2, 64), (4, 64)):
        cyc, tr = net.probe(4, n=n, iters=64, nmma=5, groups=groups, trace=True, nwarps=mode)
        t0 = tr[8, 0, 0]
        print(f"trace mode={mode} groups={groups} epilogue={'yes' if n else 'no'} nmma=5  ({cyc:.0f} cycles/iter); columns: issue, committed, epi woke, epi arrived")
        for i in range(8, 11):
            for g in range(groups):
                print(f"  iter {i} group {g}: " + "  ".join(f"{int(v - t0):7d}" for v in tr[i, g]))
