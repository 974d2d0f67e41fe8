"""RGB48 output pack with a pinned host ring and CUDA-event hand-off.

Mirror of the reference feeder pack (src/gui_pipeline_worker_feeders.py): ``_PinnedMpvFrame`` (:38-70), the ring
(:125-170) and ``_tensor_to_rgb48_bytes`` (:193-249).  The five ATen launches and two full-size device temporaries
of the reference collapse into one fused kernel (clamp -> *65535 -> +0.5 -> truncate -> CHW->HWC interleave) writing
a device staging slot that is DMA-copied into the pinned ring slot on the pack stream.

The reference applies NO transfer function here (its network output is already PQ-coded, SURVEY §8 surprise #3); that
is the default.  ``transfer="pq1000"`` is the optional linear->ST 2084 encode through a code table.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np
import torch

from . import _native


def _gpu_rgb48_ring_frames() -> int:
    try:
        value = int(str(os.environ.get("HDRTVNET_FEEDER_GPU_RGB48_RING_FRAMES", "3")).strip() or "3")
    except Exception:
        value = 3
    return max(2, min(8, value))


class PinnedFrame:
    """A CUDA-ready pinned host frame whose slot is released after the sink has written it
    (same protocol as the reference's _PinnedMpvFrame: wait_ready / buffer_view / release)."""

    def __init__(self, slot: dict, ready_event) -> None:
        self._slot = slot
        self._ready_event = ready_event
        self._ready_waited = False
        self._released = False

    def wait_ready(self) -> None:
        if self._ready_waited:
            return
        if self._ready_event is not None:
            self._ready_event.synchronize()
        self._ready_waited = True

    def buffer_view(self):
        self.wait_ready()
        return memoryview(self._slot["numpy"]).cast("B")

    def numpy(self) -> np.ndarray:
        self.wait_ready()
        return self._slot["numpy"]

    def checksum(self) -> int:
        """Descriptor checksum of the frame computed on the device by the pack kernel (process_rgb48(checksum=True));
        equals sharding.frame_checksum(self.numpy())."""
        if not self._slot.get("has_checksum"):
            raise RuntimeError("frame was produced without checksum=True")
        self.wait_ready()
        return int(self._slot["checksum"].numpy().view(np.uint64)[0])

    def release(self) -> None:
        if self._released:
            return
        self._released = True
        try:
            self.wait_ready()
        finally:
            self._slot["free"].set()


# ---- optional PQ (ST 2084) code table: half bit pattern of clamp(x,0,1) -> uint16 code ----------------
_PQ_M1, _PQ_M2 = 2610.0 / 16384.0, 2523.0 / 32.0
_PQ_C1, _PQ_C2, _PQ_C3 = 3424.0 / 4096.0, 2413.0 / 128.0, 2392.0 / 128.0


def pq_code_table(peak_nits: float = 1000.0) -> np.ndarray:
    """15361 codes for half patterns 0x0000..0x3C00 (0.0 .. 1.0).  Same formula and float32 steps as the
    reference's _linear_bgr_to_bt2100_pq_bgr_u16 (src/gui_objective_metrics.py:477-491, 531-539)."""
    x = np.arange(0x3C01, dtype=np.uint16).view(np.float16).astype(np.float32)
    lum = np.clip(x, 0.0, 1.0) * float(peak_nits)
    y = np.clip(lum.astype(np.float32, copy=False) / 10000.0, 0.0, 1.0)
    y_m1 = np.power(y, _PQ_M1).astype(np.float32, copy=False)
    num = _PQ_C1 + (_PQ_C2 * y_m1)
    den = 1.0 + (_PQ_C3 * y_m1)
    pq = np.power(num / np.maximum(den, 1e-12), _PQ_M2).astype(np.float32, copy=False)
    return np.clip((pq * 65535.0) + 0.5, 0.0, 65535.0).astype(np.uint16)


class PinnedRing:
    """Ring of pinned uint16 (H,W,3) host frames with the reference's free/claim protocol
    (gui_pipeline_worker_feeders.py:125-170).  Each slot carries a reusable CUDA event whose raw handle is stable, so
    that the C ABI can record it on one of its own streams (hdrtv_process)."""

    def __init__(self, device, ring_frames: int | None = None):
        self.device = torch.device(device)
        self.ring_frames = ring_frames or _gpu_rgb48_ring_frames()
        self._shape = None
        self._slots = []
        self._index = 0

    def _ring(self, shape):
        if self._shape == shape and self._slots:
            return self._slots
        self._slots = []
        with torch.cuda.device(self.device):
            for _ in range(self.ring_frames):
                host = torch.empty(shape, dtype=torch.uint16, pin_memory=True)
                free = threading.Event()
                free.set()
                ev = torch.cuda.Event(enable_timing=False)
                ev.record(torch.cuda.current_stream(self.device))       # materialise the handle
                self._slots.append({"tensor": host, "numpy": host.numpy(), "free": free, "event": ev, "source": None,
                                    "checksum": torch.zeros(1, dtype=torch.int64).pin_memory(), "has_checksum": False})
        self._shape = shape
        self._index = 0
        return self._slots

    def acquire(self, shape, timeout=0.25):
        slots = self._ring(shape)
        start = self._index % len(slots)
        for off in range(len(slots)):
            idx = (start + off) % len(slots)
            if slots[idx]["free"].is_set():
                slots[idx]["free"].clear()
                self._index = (idx + 1) % len(slots)
                return slots[idx]
        slot = slots[start]
        if not slot["free"].wait(timeout=timeout):
            raise RuntimeError("RGB48 ring exhausted: no pinned slot was released within "
                               f"{timeout:.2f}s (ring={len(slots)}); the consumer is not calling release()")
        slot["free"].clear()
        self._index = (start + 1) % len(slots)
        return slot


class RGB48Packer:
    """Owns the pack stream, the device staging slots and the pinned ring for one device."""

    def __init__(self, device=None, ring_frames: int | None = None, transfer: str = "identity"):
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA device not available; the B200 pack has no CPU fallback.")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("RGB48Packer needs a CUDA device (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._lib = _native.load()
        self._handle = C.c_void_p()
        cfg = _native.Config(self.device.index, _native.FP16)
        if self._lib.hdrtv_create(C.byref(cfg), C.byref(self._handle)) != 0:
            raise RuntimeError("hdrtv_create failed: " + _native.last_error(None))
        self.ring_frames = ring_frames or _gpu_rgb48_ring_frames()
        if transfer not in ("identity", "pq1000"):
            raise ValueError("transfer must be 'identity' or 'pq1000'")
        self.transfer = transfer
        if transfer == "pq1000":
            lut = np.ascontiguousarray(pq_code_table(1000.0))
            _native.check(self._lib.hdrtv_set_transfer_lut(self._handle, lut.ctypes.data, lut.size), self._handle,
                          "hdrtv_set_transfer_lut")
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.Stream()
        self._shape = None
        self._slots = []
        self._index = 0

    def _ring(self, shape):
        if self._shape == shape and self._slots:
            return self._slots
        self._slots = []
        for _ in range(self.ring_frames):
            host = torch.empty(shape, dtype=torch.uint16, pin_memory=True)
            free = threading.Event()
            free.set()
            self._slots.append({"tensor": host, "numpy": host.numpy(), "free": free,
                                "staging": torch.empty(shape, dtype=torch.uint16, device=self.device)})
        self._shape = shape
        self._index = 0
        return self._slots

    def _acquire(self, shape, timeout=0.25):
        slots = self._ring(shape)
        start = self._index % len(slots)
        for off in range(len(slots)):
            idx = (start + off) % len(slots)
            if slots[idx]["free"].is_set():
                slots[idx]["free"].clear()
                self._index = (idx + 1) % len(slots)
                return slots[idx]
        slot = slots[start]
        if not slot["free"].wait(timeout=timeout):
            raise RuntimeError("RGB48 ring exhausted: no pinned slot was released within "
                               f"{timeout:.2f}s (ring={len(slots)}); the consumer is not calling release()")
        slot["free"].clear()
        self._index = (start + 1) % len(slots)
        return slot

    def pack(self, tensor, wait_event=None) -> PinnedFrame:
        """(1,3,H,W) fp16/fp32 device tensor (or the (out, agcm) tuple) -> PinnedFrame holding uint16 (H,W,3) RGB."""
        prepared = tensor[0] if isinstance(tensor, (tuple, list)) else tensor
        if prepared.device.type != "cuda":
            raise RuntimeError("RGB48Packer.pack needs a CUDA tensor (no CPU fallback)")
        if prepared.dtype not in (torch.float16, torch.float32):
            prepared = prepared.float()
        src = prepared.contiguous()
        h, w = int(src.shape[-2]), int(src.shape[-1])
        slot = self._acquire((h, w, 3))
        dt = _native.FP16 if src.dtype == torch.float16 else _native.FP32
        tr = _native.TRANSFER_LUT if self.transfer == "pq1000" else _native.TRANSFER_IDENTITY
        producer_stream = torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            if wait_event is not None:
                self.stream.wait_event(wait_event)
            else:
                self.stream.wait_stream(producer_stream)
            _native.check(self._lib.hdrtv_pack_rgb48(self._handle, src.data_ptr(), dt, h, w, slot["staging"].data_ptr(), tr,
                                                     C.c_void_p(self.stream.cuda_stream)), self._handle, "hdrtv_pack_rgb48")
            slot["tensor"].copy_(slot["staging"], non_blocking=True)
            ready = torch.cuda.Event(enable_timing=False)
            ready.record(self.stream)
        src.record_stream(self.stream)
        return PinnedFrame(slot, ready)

    def pack_device(self, tensor, dst_u16: torch.Tensor, stream=None):
        """Pack into a caller-owned device (or mapped pinned) uint16 (H,W,3) tensor on `stream` (default: the
        current stream); no host copy, no event.  Used by device-resident pipelines and the benchmark."""
        prepared = tensor[0] if isinstance(tensor, (tuple, list)) else tensor
        src = prepared.contiguous()
        h, w = int(src.shape[-2]), int(src.shape[-1])
        dt = _native.FP16 if src.dtype == torch.float16 else _native.FP32
        tr = _native.TRANSFER_LUT if self.transfer == "pq1000" else _native.TRANSFER_IDENTITY
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        _native.check(self._lib.hdrtv_pack_rgb48(self._handle, src.data_ptr(), dt, h, w, dst_u16.data_ptr(), tr,
                                                 C.c_void_p(st.cuda_stream)), self._handle, "hdrtv_pack_rgb48")
        return dst_u16

    def launch_count(self) -> int:
        return int(self._lib.hdrtv_launch_count(self._handle))

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self._lib.hdrtv_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def tensor_to_rgb48_bytes(tensor, host_state: dict):
    """Drop-in for the reference's ``_tensor_to_rgb48_bytes(tensor, host_state)``
    (gui_pipeline_worker_feeders.py:193): returns a PinnedFrame (wait_ready / buffer_view / release).
    ``host_state`` caches the packer exactly like the reference caches its stream and staging tensors.
    The caller must have made `tensor` ready (the reference's feeder does ``ready_event.synchronize()`` first,
    :469-473) or pass ``host_state['wait_event']``."""
    prepared = tensor[0] if isinstance(tensor, (tuple, list)) else tensor
    key = str(prepared.device)
    packer = host_state.get("packer")
    if packer is None or host_state.get("packer_key") != key:
        packer = RGB48Packer(prepared.device, transfer=host_state.get("transfer", "identity"))
        host_state["packer"] = packer
        host_state["packer_key"] = key
    return packer.pack(prepared, wait_event=host_state.get("wait_event"))
