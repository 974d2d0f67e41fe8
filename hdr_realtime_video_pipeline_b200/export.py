"""Frame-sharded export of a clip to raw rgb48le (BASELINE config 4).

The reference exports serially: decode -> preprocess -> infer -> RGB48 -> ``ffmpeg stdin`` (src/gui_export.py:1034-1104,
ffmpeg command :966-1023).  Frames are independent, so here the clip is cut into contiguous chunks, one rank per GPU
(``sharding.frame_chunk``); every rank runs its own decode(or synthetic source) -> H2D -> infer -> pack -> pinned-ring
stream and writes its frames at their byte offsets of ONE raw ``rgb48le`` file (frame size is fixed, so no ordering
collective is needed on the data path).  A single gather of per-rank records (frame index, checksum) at the end proves
the file is complete and in order.  ``ffmpeg_rawvideo_args`` is the reference's own input/output contract, so the
raw file (or a pipe of it) can be handed to ``ffmpeg`` where one is installed.

No CPU fallback: the frames are produced by ``HDRTVNetB200`` / ``tensor_to_rgb48_bytes``.
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import sharding

EXPORT_HDR_TARGET_PEAK_NITS = 1000.0     # src/gui_export.py (zscale npl)


def ffmpeg_rawvideo_args(width: int, height: int, fps: float, output_path: str, source_path: str | None = None,
                         ffmpeg: str = "ffmpeg", input_path: str = "-") -> list[str]:
    """The reference's ffmpeg invocation for an rgb48le BT.2020/PQ stream -> ProRes 422 HQ 10-bit
    (src/gui_export.py:947-1023), with the raw stream on ``input_path`` (default stdin)."""
    vf = ["deband",
          "zscale=matrixin=gbr:transferin=smpte2084:primariesin=bt2020:rangein=full:matrix=bt2020nc:"
          f"transfer=smpte2084:primaries=bt2020:range=limited:dither=error_diffusion:npl={EXPORT_HDR_TARGET_PEAK_NITS:.0f}",
          "format=yuv422p10le"]
    cmd = [ffmpeg, "-y", "-hide_banner", "-loglevel", "error", "-f", "rawvideo", "-pix_fmt", "rgb48le",
           "-s:v", f"{int(width)}x{int(height)}", "-r", f"{float(fps):.6f}", "-color_range", "pc",
           "-colorspace", "bt2020nc", "-color_trc", "smpte2084", "-color_primaries", "bt2020", "-i", input_path]
    if source_path:
        cmd += ["-i", source_path, "-map", "0:v:0", "-map", "1:a?"]
    cmd += ["-vf", ",".join(vf), "-c:v", "prores_ks", "-profile:v", "3", "-pix_fmt", "yuv422p10le",
            "-bsf:v", "prores_metadata=color_primaries=bt2020:color_trc=smpte2084:colorspace=bt2020nc",
            "-color_range", "tv", "-colorspace", "bt2020nc", "-color_trc", "smpte2084", "-color_primaries", "bt2020",
            "-vendor", "apl0"]
    if source_path:
        cmd += ["-c:a", "pcm_s16le"]
    cmd += ["-movflags", "+faststart+write_colr", output_path]
    return cmd


class Rgb48RawWriter:
    """Ordered raw rgb48le file shared by all ranks: frame i lives at byte offset i * H * W * 6."""

    def __init__(self, path: str, n_frames: int, height: int, width: int, create: bool):
        self.path, self.n_frames = path, int(n_frames)
        self.frame_bytes = int(height) * int(width) * 6
        if create:
            with open(path, "wb") as f:
                f.truncate(self.n_frames * self.frame_bytes)
        self._fd = os.open(path, os.O_WRONLY)

    def write(self, frame_idx: int, payload) -> None:
        if not (0 <= frame_idx < self.n_frames):
            raise ValueError(f"frame index {frame_idx} outside the clip (0..{self.n_frames - 1})")
        view = memoryview(payload).cast("B") if not isinstance(payload, memoryview) else payload
        if view.nbytes != self.frame_bytes:
            raise ValueError(f"frame {frame_idx}: {view.nbytes} bytes, expected {self.frame_bytes}")
        done = 0
        while done < view.nbytes:
            done += os.pwrite(self._fd, view[done:], frame_idx * self.frame_bytes + done)

    def close(self) -> None:
        if self._fd is not None:
            os.close(self._fd)
            self._fd = None


def export_clip(processor, frames, n_frames: int, out_path: str | None, rank: int = 0, world_size: int = 1, pack=None,
                barrier=None, one_call: bool = True, in_flight: int = 3, device_checksums: bool = True) -> dict:
    """Export this rank's contiguous chunk of an ``n_frames`` clip.

    processor : an ``HDRTVNetB200`` (anything with preprocess / infer)
    frames    : callable ``frames(i) -> uint8 HxWx3 BGR`` (decode or synthetic source), called only for this rank's frames
    pack      : ``pack(out) -> payload`` with wait_ready()/buffer_view()/release(); defaults to tensor_to_rgb48_bytes
    barrier   : optional callable run after rank 0 created the file and before anyone writes (dist.barrier)
    one_call  : with the default pack, use ``processor.process_rgb48`` (one C-ABI call per frame, ``hdrtv_process``: the
                frame is DMA-copied straight from the array ``frames(i)`` returned, which must not be rewritten before
                the frame has been written out) instead of preprocess -> infer -> pack; the bytes are identical
    out_path  : raw rgb48le file shared by all ranks, or None: the frames are delivered to the pinned ring and released
                (the sink of the benchmark: "decode-to-RGB48 output", BASELINE config 3 / 4 without a disk in the loop)
    in_flight : frames submitted ahead of the one being drained (the pinned ring is sized in_flight + 1)
    device_checksums : (one-call path) the descriptor checksum of every frame is computed by the pack kernel on the GPU
                (``hdrtv_process_ex``) instead of a numpy pass over the 50 MB frame on the host
    Returns the per-rank record; gather with ``sharding.gather_run_records`` and check with ``sharding.merge_descriptors``.
    """
    from .feeders import tensor_to_rgb48_bytes

    first, last = sharding.frame_chunk(n_frames, rank, world_size)
    probe = frames(first) if last > first else None
    state: dict = {}
    use_one_call = bool(one_call) and pack is None and hasattr(processor, "process_rgb48")
    if pack is None:
        def pack(out):
            return tensor_to_rgb48_bytes(out, state)
    h, w = (probe.shape[0], probe.shape[1]) if probe is not None else (0, 0)
    writer = None
    if out_path is not None:
        if rank == 0:                                     # rank 0 sizes the file; it always owns frame 0 of a non-empty clip
            if probe is None:
                raise ValueError("empty clip")
            writer = Rgb48RawWriter(out_path, n_frames, h, w, create=True)
        if barrier is not None:
            barrier()
        if writer is None and probe is not None:
            writer = Rgb48RawWriter(out_path, n_frames, h, w, create=False)
    elif barrier is not None:
        barrier()
    in_flight = max(1, int(in_flight))
    if use_one_call and hasattr(processor, "set_rgb48_ring_frames"):
        processor.set_rgb48_ring_frames(in_flight + 1)
    dev_cks = bool(device_checksums) and use_one_call
    descriptors, pending = [], []
    t0 = time.perf_counter()

    def drain(entry):
        idx, payload = entry
        view = payload.buffer_view()                      # waits for the CUDA event of the ring slot
        cks = payload.checksum() if dev_cks else sharding.frame_checksum(np.frombuffer(view, dtype=np.uint16))
        descriptors.append((idx, cks))
        if writer is not None:
            writer.write(idx, view)
        payload.release()

    for i in range(first, last):
        frame = probe if i == first else frames(i)
        if use_one_call:
            pending.append((i, processor.process_rgb48(frame, checksum=dev_cks)))
        else:
            pending.append((i, pack(processor.infer(processor.preprocess(frame)))))
        if len(pending) > in_flight or (not use_one_call and len(pending) >= 2):
            drain(pending.pop(0))
    for entry in pending:
        drain(entry)
    elapsed = time.perf_counter() - t0
    if writer is not None:
        writer.close()
    return {"rank": rank, "first_frame": first, "n_frames": last - first, "elapsed_s": elapsed, "descriptors": descriptors,
            "height": h, "width": w}
