"""B200-native HDRTVNet++ per-frame SDR->HDR inference (AGCM + LE, BGR24 in -> RGB48 out).

Drop-in for the reference's model wrapper and feeder pack on this path only:
    HDRTVNetB200            <-> HDRTVNetTorch / HDRTVNetTensorRT   (src/models/hdrtvnet_torch.py)
    tensor_to_rgb48_bytes   <-> _tensor_to_rgb48_bytes             (src/gui_pipeline_worker_feeders.py)
"""
from .backend import HDRTVNetB200, load_state_dict_any  # noqa: F401
from .feeders import PinnedFrame, RGB48Packer, pq_code_table, tensor_to_rgb48_bytes  # noqa: F401
from .export import Rgb48RawWriter, export_clip, ffmpeg_rawvideo_args  # noqa: F401
from .sharding import frame_chunk, gather_run_records  # noqa: F401
from .synth import synth_clip, synth_frame  # noqa: F401

__all__ = ["HDRTVNetB200", "load_state_dict_any", "PinnedFrame", "RGB48Packer", "pq_code_table",
           "tensor_to_rgb48_bytes", "Rgb48RawWriter", "export_clip", "ffmpeg_rawvideo_args", "frame_chunk", "gather_run_records", "synth_clip", "synth_frame"]
