"""Frame-sharded multi-GPU layout: frames are independent, so a clip is cut into contiguous chunks,
one process per GPU, no collective on the per-pixel path.  One gather at the end of a run brings the per-rank
metrics and the ordered output descriptors (frame index, checksum) to every rank.

The reference has no distributed code at all (single process, cuda:0 — hdrtvnet_torch.py:1682); its serial export
loop (src/gui_export.py:1072-1104) is what gets partitioned here.
"""
from __future__ import annotations

import numpy as np


def frame_chunk(n_frames: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous chunk [first, last) of rank `rank`: floor(r*N/G) .. floor((r+1)*N/G)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    if n_frames < 0:
        raise ValueError("n_frames must be >= 0")
    return (rank * n_frames) // world_size, ((rank + 1) * n_frames) // world_size


def pin_rank_to_local_cores(local_rank: int, local_world: int, device_of_rank=None) -> list[int]:
    """One process per GPU: bind this rank to its own share of the host cores BEFORE it allocates pinned frame buffers
    (first touch places them on the NUMA node of the allocating core).  NVML names the cores close to each GPU; ranks
    whose GPUs share a core set split that set evenly; without NVML the allowed cores are split evenly over the local
    ranks.  Returns the cores this process now runs on.  (torchrun starts every rank on the same affinity mask: eight
    submit loops and all their pinned memory on one NUMA node cost 16 % of the 8-GPU end-to-end throughput in round 1.)"""
    import os
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return []
    allowed_set = set(allowed)
    dev = device_of_rank or (lambda r: r)
    near = {}
    try:
        import pynvml
        pynvml.nvmlInit()
        for r in range(local_world):
            h = pynvml.nvmlDeviceGetHandleByIndex(int(dev(r)))
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(allowed) // 64) + 1)
            cores = tuple(c for c in (64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1)
                          if c in allowed_set)
            near[r] = cores
    except Exception:
        near = {}
    mine = None
    if near.get(local_rank):
        peers = [r for r in range(local_world) if near.get(r) == near[local_rank]]
        cores = list(near[local_rank])
        if len(cores) >= len(peers):
            share = len(cores) // len(peers)
            k = peers.index(local_rank)
            mine = cores[k * share:(k + 1) * share]
    if not mine:
        share = max(1, len(allowed) // max(1, local_world))
        mine = allowed[local_rank * share:(local_rank + 1) * share] or allowed
    try:
        os.sched_setaffinity(0, set(mine))
    except OSError:
        return allowed
    # torch sized its intra-op pool for the affinity mask the process STARTED with (all cores of the box): on this rank's
    # share that pool oversubscribes the cores (8 ranks x 32 threads on 32 cores) and every host-side tensor copy of the
    # frame loop (the staging copy into pinned memory of `preprocess`) fights for them
    try:
        import torch
        torch.set_num_threads(max(1, len(mine)))
    except Exception:
        pass
    return sorted(mine)


def frame_checksum(rgb48: np.ndarray) -> int:
    """Order-sensitive 64-bit checksum of one packed frame (descriptor payload, cheap to compare across runs)."""
    a = np.ascontiguousarray(rgb48).view(np.uint16).astype(np.uint64).ravel()
    idx = (np.arange(a.size, dtype=np.uint64) % np.uint64(65521)) + np.uint64(1)
    return int((a * idx).sum(dtype=np.uint64))


def gather_run_records(record: dict, group=None) -> list[dict]:
    """all_gather of one small python record per rank (NCCL on the GPU box, gloo in CPU tests).
    Returns the records ordered by rank; with no initialised process group returns [record]."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return [record]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, record, group=group)
    return out


def merge_descriptors(records: list[dict]) -> list[tuple[int, int]]:
    """Ordered (frame_idx, checksum) list of the whole clip from the per-rank records; checks the chunks tile the
    clip exactly once and in order."""
    merged = []
    expect = None
    for r in sorted(records, key=lambda r: r["first_frame"]):
        if expect is not None and r["first_frame"] != expect:
            raise ValueError(f"frame chunks do not tile the clip: expected {expect}, got {r['first_frame']}")
        desc = list(r.get("descriptors", []))
        if desc and [d[0] for d in desc] != list(range(r["first_frame"], r["first_frame"] + r["n_frames"])):
            raise ValueError("descriptor frame indices out of order")
        merged.extend(desc)
        expect = r["first_frame"] + r["n_frames"]
    return merged
