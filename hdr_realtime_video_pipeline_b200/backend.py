"""HDRTVNetB200 — drop-in for the reference's model wrapper on the per-frame SDR->HDR path.

Mirrors ``HDRTVNetTorch`` / ``HDRTVNetTensorRT`` (reference src/models/hdrtvnet_torch.py:1513-2472, :8164-9106):
same constructor keywords, same ``preprocess / infer / postprocess / process / process_timed / warmup_compile /
end_profiling`` methods, same attributes the callers read (gui_pipeline_worker_model.py, gui_export.py, main.py,
cli_playback_benchmark.py).  Underneath it calls the sm_100a kernels through the C ABI in
``include/hdrtv_b200.h``.  PyTorch is used for device memory, streams and events only.

There is no CPU path, no TensorRT, no Triton and no eager fallback: failures raise.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np
import torch

from . import _native

_VALID_PRECISIONS = {"auto", "fp16", "fp32", "int8-full", "int8-mixed"}
_HERE = os.path.dirname(os.path.abspath(__file__))


def _env_bool(name: str, default: bool) -> bool:
    v = os.environ.get(name)
    if v is None:
        return default
    return str(v).strip().lower() in {"1", "true", "yes", "on"}


def _assume_aligned_shapes_for_resolution(width: int, height: int) -> bool:
    # hdrtvnet_torch.py:204-207: only the two GUI presets are declared aligned.
    return (int(width), int(height)) in {(1920, 1080), (1280, 720)}


def load_state_dict_any(model_path):
    """Accepts a dict, a ``.npz`` of fp32 arrays, or a torch checkpoint: raw state-dict (HR.pt) or a
    ``{"state_dict": ..., "architecture": ...}`` source checkpoint (hdrtvnet_torch.py:1491-1498)."""
    arch = {}
    if isinstance(model_path, dict):
        payload = model_path
    else:
        ext = os.path.splitext(str(model_path))[1].lower()
        if not os.path.isfile(model_path):
            raise FileNotFoundError(f"model weights not found: {model_path}")
        if ext == ".npz":
            with np.load(model_path) as z:
                payload = {k: z[k] for k in z.files}
        else:
            payload = torch.load(model_path, map_location="cpu", weights_only=True)
    if isinstance(payload, dict) and "state_dict" in payload:
        a = payload.get("architecture") or {}
        arch = a if isinstance(a, dict) else {}
        payload = payload.get("state_dict") or {}
    state = {}
    for k, v in payload.items():
        k = k[7:] if k.startswith("module.") else k          # hdrtvnet_torch.py:2154-2157
        if isinstance(v, torch.Tensor):
            v = v.detach().to(torch.float32).cpu().numpy()
        state[k] = np.ascontiguousarray(np.asarray(v, dtype=np.float32))
    return state, arch


def split_int8_state(state: dict):
    """Eager INT8 checkpoint (hdrtvnet_torch.py:1748-1963; layers W8Conv2d / W8A8Conv2d / W8A8Linear, :233-410) ->
    (ordinary fp32 state-dict with de-quantised weights, {layer: (x_scale, x_zero, mode)}).

    weight = weight_int8 * w_scale per output channel (:361); a layer with ``x_scale`` fake-quantises its input,
    asymmetrically when it also has ``x_zero`` (:350-360).  Layers stored as plain ``weight`` pass through.
    ``raw8`` (third result): {"<layer>.weight_int8": int8 weights as fp32, "<layer>.w_scale": scales} of the W8A8 layers, for
    the kind::i8 tensor-core kernels."""
    out, quant, raw8 = {}, {}, {}
    for k, v in state.items():
        if k.endswith(".weight_int8"):
            layer = k[: -len(".weight_int8")]
            # W8A8 layers carry ".w_scale" (:361), weight-only W8 layers ".scale" (W8Conv2d / W8Linear, :233-291)
            sc = state[layer + ".w_scale"] if (layer + ".w_scale") in state else state[layer + ".scale"]
            scale = np.asarray(sc, dtype=np.float32).reshape((-1,) + (1,) * (v.ndim - 1))
            out[layer + ".weight"] = np.ascontiguousarray(np.asarray(v, dtype=np.float32) * scale)
            if (layer + ".x_scale") in state:
                raw8[layer + ".weight_int8"] = np.ascontiguousarray(np.asarray(v, dtype=np.float32))
                raw8[layer + ".w_scale"] = np.ascontiguousarray(np.asarray(sc, dtype=np.float32).reshape(-1))
        elif k.endswith(".x_scale"):
            layer = k[: -len(".x_scale")]
            zero = state.get(layer + ".x_zero")
            quant[layer] = (float(np.asarray(v).reshape(-1)[0]), float(np.asarray(zero).reshape(-1)[0]) if zero is not None else 0.0,
                            2 if zero is not None else 1)
        elif k.endswith(".w_scale") or k.endswith(".x_zero") or (k.endswith(".scale") and (k[: -len(".scale")] + ".weight_int8") in state):
            continue
        else:
            out[k] = v
    return out, quant, raw8


# The W8A8 layers of the reference's shipping INT8 layout, "INT8 Mixed QAT" (configs/qat_layouts/original_nohg_mixed_w8a8.txt):
# the layers whose input quantisers the tensor-core path implements (kind::i8 launches for the stand-alone 3x3 convs, f16 MMAs
# on values de-quantised by the producer for the ones inside fused kernels).
MIXED_W8A8_LAYERS = frozenset(
    ["LE.down_conv1", "LE.down_conv2", "LE.down_conv3", "LE.up_conv1.0", "LE.up_conv2.0", "LE.up_conv3.0",
     "LE.recon_trunk1.0.conv1", "LE.recon_trunk1.0.conv2", "LE.recon_trunk2.0.conv1", "LE.recon_trunk2.0.conv2",
     "LE.recon_trunk4.0.conv1", "LE.recon_trunk4.0.conv2", "LE.recon_trunk5.0.conv2",
     "LE.CondNet1.4", "LE.CondNet2.4", "LE.CondNet3.0", "LE.CondNet3.2", "LE.CondNet3.4", "LE.CondNet4.0", "LE.CondNet4.2",
     "LE.CondNet4.4"] + [f"LE.recon_trunk3.{j}.conv{k}" for j in range(4) for k in (1, 2)])


def is_int8_state(state: dict) -> bool:
    return any(k.endswith(".weight_int8") for k in state)


class HDRTVNetB200:
    """B200-native backend with the reference wrapper's interface."""

    def __init__(self, model_path, device="auto", precision="auto",
                 compile_model=True, force_compile=False, compile_mode="auto",
                 use_cuda_graphs=False, force_channels_last=False,
                 predequantize="auto", hg_weights=None, use_hg=True,
                 warmup_passes=3, fast_condition_resize=False,
                 # HDRTVNetTensorRT extras (hdrtvnet_torch.py:8172-8189): accepted and ignored
                 engine_width=None, engine_height=None, mode_name=None, qdq_fusion=None, keep_onnx=None,
                 debug_library=False, **ignored_tensorrt_kwargs):
        self.model_path = model_path
        self._warmup_passes = warmup_passes
        self._fast_condition_resize = bool(fast_condition_resize) or _env_bool("HDRTVNET_FAST_COND_RESIZE", False)
        self._fast_zero_condition = _env_bool("HDRTVNET_ZERO_COND", False)
        # argument validation first (ValueError), availability second (RuntimeError) — hdrtvnet_torch.py:1678-1702
        if str(device).lower() not in ("auto", "cuda", "cpu") and not str(device).lower().startswith("cuda:"):
            raise ValueError("device must be one of: auto, cuda, cpu")
        if str(precision).lower() not in _VALID_PRECISIONS:
            raise ValueError("precision must be one of: auto, fp16, fp32, int8-full, int8-mixed")
        self.device = self._resolve_device(device)
        self.precision = self._resolve_precision(precision, self.device)
        self._use_cuda = True
        # INT8 layouts: the reference's eager INT8 model is fake-quantisation around ordinary convolutions
        # (hdrtvnet_torch.py:350-364); here it runs on the FP32 CUDA-core path with the same quantisers.
        self._int8 = self.precision in ("int8-full", "int8-mixed")
        state, arch, quant, raw8 = self._read_checkpoint(model_path)
        # INT8 checkpoints whose quantised layers are those of the mixed layout run on the FP16 tensor-core path with
        # tcgen05.mma.kind::i8 launches for the W8A8 convs; any other layout (Full-QAT: all 128 layers) runs the reference's
        # fake-quantisation arithmetic on the FP32 CUDA-core path.  HDRTV_B200_INT8_FP32=1 forces the latter.
        self._int8_tensor_path = bool(quant) and set(quant) <= MIXED_W8A8_LAYERS and all(q[2] == 2 for q in quant.values()) \
            and not _env_bool("HDRTV_B200_INT8_FP32", False)
        half = self.precision == "fp16" or self._int8_tensor_path
        self._dtype = torch.float16 if half else torch.float32
        self._np_dtype = np.float16 if half else np.float32
        self._hg_weights_explicit = hg_weights is not None
        self._hg_weights = hg_weights
        self._use_hg = bool(use_hg)
        self._is_flat_model = False
        self._is_w8_model = False
        self._compiled = False
        self._compile_mode = None
        self._memory_format_name = "contiguous"
        self._use_channels_last = False
        self._assume_aligned_shapes = None
        self._trt_engine = None
        self._trt_context = None
        self.engine_path = None
        self.model = None            # like the TensorRT wrapper after build (hdrtvnet_torch.py:8429)
        self.expected_hw = None
        self.is_static_input_model = False

        # debug_library=True (tests / scripts): the test build of the engine, which adds the debug, self-test and probe
        # entry points of include/hdrtv_b200_test.h to the product ABI
        self._debug_library = bool(debug_library) or _env_bool("HDRTV_B200_TEST_LIB", False)
        self._lib = _native.load_test() if self._debug_library else _native.load()
        self._handle = C.c_void_p()
        cfg = _native.Config(self.device.index, _native.FP16 if half else _native.FP32)
        self._is_w8_model = self._int8
        rc = self._lib.hdrtv_create(C.byref(cfg), C.byref(self._handle))
        if rc != 0:
            raise RuntimeError("hdrtv_create failed: " + _native.last_error(None, self._lib))
        self._load_model(model_path, state, arch, quant, raw8)

        # Frame pipelining: preprocess() puts the H2D copy, the normalise / condition kernels and the AGCM condition
        # classifier (which depends on `cond` alone) on a side stream, so they overlap the previous frame's LE network;
        # infer() then skips the classifier.  Results are identical; HDRTV_B200_PIPELINE=0 keeps everything in-stream.
        self._pipeline = _env_bool("HDRTV_B200_PIPELINE", True)
        self._side = None
        self._rgb48_ring = None
        self._rgb48_ring_frames = None     # None: the reference's default / HDRTVNET_FEEDER_GPU_RGB48_RING_FRAMES
        self._lut_set = False
        self._proc_dirty = False          # hdrtv_process frames in flight on the context's own streams
        self._ev_inputs_free = self._ev_pre_done = None
        self._cls_ready = False
        self._buf_hw = None
        self._gpu_input = self._gpu_cond = self._gpu_raw = None
        self._gpu_out = self._gpu_agcm = self._gpu_u8 = None
        self._gpu_hg_out = None
        self._pin_input = self._pin_output = None
        print(f"GPU: {torch.cuda.get_device_name(self.device)} (CUDA, sm_100a kernels)")
        print(f"B200 device : {self.device}")
        print(f"B200 precision: {self.precision}")
        if self._warmup_passes and self._warmup_passes > 0:
            self._warmup()

    # ------------------------------------------------------------------ device / precision (hdrtvnet_torch.py:1678-1702)
    def _resolve_device(self, device):
        mode = str(device).lower()
        if mode in ("auto", "cuda") or mode.startswith("cuda:"):
            if not torch.cuda.is_available():
                raise RuntimeError("CUDA device not available; the B200 backend has no CPU fallback.")
            idx = int(mode.split(":", 1)[1]) if mode.startswith("cuda:") else torch.cuda.current_device()
            return torch.device("cuda", idx)
        if mode == "cpu":
            raise RuntimeError("device='cpu' is not supported: the B200 backend has no CPU fallback.")
        raise ValueError("device must be one of: auto, cuda, cpu")

    def _resolve_precision(self, precision, device):
        p = str(precision).lower()
        if p not in _VALID_PRECISIONS:
            raise ValueError("precision must be one of: auto, fp16, fp32, int8-full, int8-mixed")
        return "fp16" if p == "auto" else p

    # ------------------------------------------------------------------ weights (hdrtvnet_torch.py:2044-2169)
    def _read_checkpoint(self, model_path):
        state, arch = load_state_dict_any(model_path)
        quant, raw8 = {}, {}
        if is_int8_state(state):
            if not self._int8:
                raise ValueError(f"{model_path} is an INT8 checkpoint; use precision='int8-full' or 'int8-mixed'")
            state, quant, raw8 = split_int8_state(state)
        elif self._int8:
            raise ValueError(f"{model_path} is not an INT8 checkpoint.\n"
                             "  Re-run: python scripts/quantize/quantize_int8_full.py or "
                             "python scripts/quantize/quantize_int8_mixed.py")      # hdrtvnet_torch.py:1758-1762
        return state, arch, quant, raw8

    def _load_model(self, model_path, state, arch, quant, raw8):
        self._act_quant = quant
        classifier = str(arch.get("classifier", os.environ.get("HDRTVNET_CLASSIFIER", "color_condition"))).strip()
        le_arch = str(arch.get("le_arch", os.environ.get("HDRTVNET_LE_ARCH", "")) or "").strip()
        post = str(arch.get("post_correction", os.environ.get("HDRTVNET_POST_CORRECTION", "")) or "").strip()
        if (classifier or "color_condition") != "color_condition" or le_arch not in ("", "None", "sft") or post not in ("", "None"):
            raise RuntimeError(
                f"unsupported architecture (classifier={classifier!r}, le_arch={le_arch!r}, post_correction={post!r}); "
                "this backend implements classifier='color_condition', le_arch='sft' (HDRUNet3T1) only")
        self._hg_state = None
        if self._use_hg:
            hg = self._resolve_hg_weights(model_path)
            if hg is not None and self._int8:
                print("WARNING: the INT8 layouts of the B200 backend cover AGCM+LE only (the reference's HG INT8 checkpoints "
                      "are absent from its tree); continuing with the no-HG model.")
                self._use_hg = False
            elif hg is not None:
                # HG_Composite (hdrtvnet_torch.py:2121-2143): base model + Hallucination_Generator, strict key check in the engine
                hg_state, hg_ckpt_arch = load_state_dict_any(hg)
                hg_arch = str(arch.get("hg_arch", hg_ckpt_arch.get("hg_arch", os.environ.get("HDRTVNET_HG_ARCH", ""))) or "").strip().lower()
                if hg_arch.replace("-", "").replace("_", "") not in ("", "none", "pixelshuffle", "fusedbn"):
                    raise RuntimeError(f"unsupported hg_arch {hg_arch!r}; this backend implements the PixelShuffle "
                                       "Hallucination_Generator (with or without the FusedBN fold)")
                self._hg_state = {k: v for k, v in hg_state.items() if not k.endswith("num_batches_tracked")}
                self._hg_weights = hg if not isinstance(hg, dict) else "<state-dict>"
            elif self._hg_weights_explicit:
                raise FileNotFoundError(f"HG weights not found: {self._hg_weights}\n"
                                        "  Check --hg-weights or disable HG with --use-hg 0.")
            else:
                print("WARNING: HG weights not found; continuing with no-HG model.")
                self._use_hg = False
        if quant:          # before the weights: on the tensor-core path the weight pack builds the kind::i8 operands from them
            names = sorted(quant)
            arr_n = (C.c_char_p * len(names))(*[n.encode() for n in names])
            arr_s = (C.c_float * len(names))(*[quant[n][0] for n in names])
            arr_z = (C.c_float * len(names))(*[quant[n][1] for n in names])
            arr_m = (C.c_int * len(names))(*[quant[n][2] for n in names])
            self._check(self._lib.hdrtv_set_act_quant(self._handle, arr_n, arr_s, arr_z, arr_m, len(names)), "hdrtv_set_act_quant")
        tensors = dict(state)
        if self._int8_tensor_path:
            tensors.update(raw8)
        descs = (_native.TensorDesc * len(tensors))()
        keep = []
        for i, (k, v) in enumerate(tensors.items()):
            v = np.ascontiguousarray(v, dtype=np.float32)
            keep.append(v)
            descs[i].name = k.encode()
            descs[i].data = v.ctypes.data_as(C.POINTER(C.c_float))
            descs[i].ndim = v.ndim
            for d in range(v.ndim):
                descs[i].shape[d] = v.shape[d]
        self._check(self._lib.hdrtv_set_weights(self._handle, descs, len(tensors)), "hdrtv_set_weights")
        if self._hg_state is not None:
            self.set_hg_weights(self._hg_state)
            self._hg_state = None
        self._n_params = int(sum(v.size for v in state.values()))
        self._layer_shapes = {k[: -len(".weight")]: tuple(v.shape) for k, v in state.items() if k.endswith(".weight")}

    def _check(self, rc, what):
        _native.check(rc, self._handle, what, self._lib)

    def set_hg_weights(self, hg_state):
        """Install (dict of arrays) or remove (None) the HG stage: model.hg.load_state_dict(strict=True),
        hdrtvnet_torch.py:2141-2143.  The engine folds eval-mode BatchNorm and repacks the weights."""
        if hg_state is None:
            self._check(self._lib.hdrtv_set_hg_weights(self._handle, None, 0), "hdrtv_set_hg_weights")
            self._use_hg = False
            return
        items = [(k, np.ascontiguousarray(v, dtype=np.float32)) for k, v in hg_state.items() if not k.endswith("num_batches_tracked")]
        descs = (_native.TensorDesc * len(items))()
        for i, (k, v) in enumerate(items):
            if v.ndim > 4:
                raise ValueError(f"HG tensor {k} has {v.ndim} dimensions")
            descs[i].name = k.encode()
            descs[i].data = v.ctypes.data_as(C.POINTER(C.c_float))
            descs[i].ndim = v.ndim
            for d in range(v.ndim):
                descs[i].shape[d] = v.shape[d]
        self._check(self._lib.hdrtv_set_hg_weights(self._handle, descs, len(items)), "hdrtv_set_hg_weights")
        self._use_hg = True

    def _need_debug_library(self):
        if not self._debug_library:
            raise RuntimeError("this call needs the test build of the engine: HDRTVNetB200(..., debug_library=True)")

    def _resolve_hg_weights(self, model_path):
        if isinstance(self._hg_weights, dict):        # extension: an in-memory state-dict (tests, benchmarks with seeded weights)
            return self._hg_weights
        cands = [self._hg_weights]
        if not isinstance(model_path, dict):
            cands.append(os.path.join(os.path.dirname(os.path.abspath(str(model_path))), "HG.pt"))
        cands.append(os.path.join(os.getcwd(), "src", "models", "weights", "original", "HG.pt"))
        for p in cands:
            if p and os.path.isfile(os.path.abspath(os.path.expanduser(str(p)))):
                return p
        return None

    def _configure_assume_aligned_shapes(self, width: int, height: int) -> None:
        # The kernels handle aligned and centre-cropped skips alike; the flag is kept for callers that set it.
        self._assume_aligned_shapes = _assume_aligned_shapes_for_resolution(width, height)

    # ------------------------------------------------------------------ buffers (hdrtvnet_torch.py:2198-2233)
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ensure_buffers(self, h, w):
        self._configure_assume_aligned_shapes(w, h)
        if self._buf_hw == (h, w):
            return
        if h < 16 or w < 16:
            raise ValueError("frames must be at least 16x16")
        torch.cuda.synchronize(self.device)
        self._check(self._lib.hdrtv_prepare(self._handle, h, w), "hdrtv_prepare")
        self._buf_hw = (h, w)
        ch, cw = max(1, h // 4), max(1, w // 4)
        dev, dt = self.device, self._dtype
        self._gpu_input = torch.empty((1, 3, h, w), dtype=dt, device=dev)
        self._gpu_cond = torch.empty((1, 3, ch, cw), dtype=dt, device=dev)
        self._gpu_out = torch.empty((1, 3, h, w), dtype=dt, device=dev)
        self._gpu_agcm = torch.empty((1, 3, h, w), dtype=dt, device=dev)
        # HG_Composite returns a float32 tensor in both precisions (mask.float() * out + img promotes)
        self._gpu_hg_out = torch.empty((1, 3, h, w), dtype=torch.float32, device=dev) if self._use_hg else None
        self._gpu_raw = torch.empty((h, w, 3), dtype=torch.uint8, device=dev)
        self._gpu_u8 = torch.empty((h, w, 3), dtype=torch.uint8, device=dev)
        self._pin_inputs = [torch.empty((h, w, 3), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        self._pin_events = [None, None]          # H2D of the slot's previous frame
        self._pin_idx = 0
        self._pin_input = self._pin_inputs[0]
        self._pin_output = torch.empty((h, w, 3), dtype=torch.uint8, pin_memory=True)
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        self._ev_inputs_free = torch.cuda.Event()
        self._ev_pre_done = torch.cuda.Event()
        self._ev_inputs_free.record(torch.cuda.current_stream(dev))     # materialise the handles
        self._ev_pre_done.record(torch.cuda.current_stream(dev))
        self._cls_ready = False

    def _leave_process_mode(self):
        """preprocess()/infer() after process_rgb48(): order them behind the one-call path's internal streams."""
        if self._proc_dirty:
            cur = torch.cuda.current_stream(self.device)
            self._check(self._lib.hdrtv_process_flush(self._handle, C.c_void_p(cur.cuda_stream)), "hdrtv_process_flush")
            if self._side is not None:
                self._side.wait_stream(cur)
            self._proc_dirty = False

    # ------------------------------------------------------------------ preprocess (hdrtvnet_torch.py:2239-2296)
    def _launch_preprocess(self, raw_dev, h, w, stream):
        # hdrtvnet_torch.py:2265-2294: zero condition > bilinear (fast_condition_resize) > antialiased bicubic
        mode = (_native.COND_ZERO if self._fast_zero_condition else
                (_native.COND_BILINEAR if self._fast_condition_resize else _native.COND_BICUBIC_AA))
        sp = C.c_void_p(stream.cuda_stream)
        if self._pipeline:       # fused front end: normalise (+ tensor-core staging) + condition image + AGCM classifier
            self._check(self._lib.hdrtv_preprocess_classify(self._handle, raw_dev.data_ptr(), h, w, self._gpu_input.data_ptr(),
                                                            self._gpu_cond.data_ptr(), mode, sp), "hdrtv_preprocess_classify")
        else:
            self._check(self._lib.hdrtv_preprocess(self._handle, raw_dev.data_ptr(), h, w, self._gpu_input.data_ptr(),
                                                   self._gpu_cond.data_ptr(), mode, sp), "hdrtv_preprocess")

    @torch.inference_mode()
    def preprocess(self, frame_bgr):
        if not isinstance(frame_bgr, np.ndarray) or frame_bgr.dtype != np.uint8 or frame_bgr.ndim != 3 or frame_bgr.shape[2] != 3:
            raise ValueError("frame_bgr must be a uint8 HxWx3 BGR array")
        h, w = frame_bgr.shape[:2]
        with torch.cuda.device(self.device):
            self._ensure_buffers(h, w)
            self._leave_process_mode()
            slot = self._pin_idx
            self._pin_idx ^= 1
            if self._pin_events[slot] is not None:
                self._pin_events[slot].synchronize()              # the slot's previous H2D copy has left the pinned buffer
            pin = self._pin_inputs[slot]
            self._pin_input = pin
            pin.copy_(torch.from_numpy(np.ascontiguousarray(frame_bgr)))
            cur = torch.cuda.current_stream(self.device)
            work = self._side if self._pipeline else cur
            if self._pipeline:
                work.wait_event(self._ev_inputs_free)             # previous infer() has read x / cond / folded weights
            with torch.cuda.stream(work):
                self._gpu_raw.copy_(pin, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(work)
                self._pin_events[slot] = ev
                self._launch_preprocess(self._gpu_raw, h, w, work)
                if self._pipeline:
                    self._ev_pre_done.record(work)
            if self._pipeline:
                cur.wait_event(self._ev_pre_done)
                self._cls_ready = True
        return self._gpu_input, self._gpu_cond

    def preprocess_device(self, frame_u8_dev: torch.Tensor, assume_ready: bool = False):
        """Extension: frame already resident on the device as uint8 (H,W,3) BGR (decode-on-GPU callers, benchmarks).
        assume_ready=True: the frame's producer finished long ago (static test frames) — the side stream does not
        wait for the current stream, so the work overlaps the previous frame's network."""
        h, w = int(frame_u8_dev.shape[0]), int(frame_u8_dev.shape[1])
        with torch.cuda.device(self.device):
            self._ensure_buffers(h, w)
            self._leave_process_mode()
            src = frame_u8_dev.contiguous()
            cur = torch.cuda.current_stream(self.device)
            work = self._side if self._pipeline else cur
            if self._pipeline:
                if not assume_ready:
                    work.wait_stream(cur)
                work.wait_event(self._ev_inputs_free)
            with torch.cuda.stream(work):
                self._launch_preprocess(src, h, w, work)
                if self._pipeline:
                    self._ev_pre_done.record(work)
            if self._pipeline:
                src.record_stream(work)
                cur.wait_event(self._ev_pre_done)
                self._cls_ready = True
        return self._gpu_input, self._gpu_cond

    # ------------------------------------------------------------------ infer (hdrtvnet_torch.py:2302-2346)
    @torch.inference_mode()
    def infer(self, input_cond):
        tensor, cond = input_cond
        if tensor.dim() != 4 or tensor.shape[0] != 1 or tensor.shape[1] != 3:
            raise ValueError("infer expects a (1,3,H,W) tensor")
        h, w = int(tensor.shape[2]), int(tensor.shape[3])
        with torch.cuda.device(self.device):
            self._ensure_buffers(h, w)
            self._leave_process_mode()
            t = tensor.to(device=self.device, dtype=self._dtype).contiguous()
            c = cond.to(device=self.device, dtype=self._dtype).contiguous()
            if tuple(c.shape) != (1, 3, max(1, h // 4), max(1, w // 4)):
                raise ValueError(f"condition tensor must be (1,3,{h // 4},{w // 4}), got {tuple(c.shape)}")
            # classifier already run by preprocess() on the side stream?  Only for exactly the tensors it produced.
            skip = bool(self._pipeline and self._cls_ready and t.data_ptr() == self._gpu_input.data_ptr()
                        and c.data_ptr() == self._gpu_cond.data_ptr())
            self._cls_ready = False
            if self._pipeline:      # any side-stream work of the last preprocess() is ordered before this infer
                torch.cuda.current_stream(self.device).wait_event(self._ev_pre_done)
            rc = self._lib.hdrtv_infer_ex(self._handle, t.data_ptr(), c.data_ptr(), h, w, self._gpu_out.data_ptr(),
                                          self._gpu_agcm.data_ptr(), 1 if skip else 0,
                                          C.c_void_p(self._ev_inputs_free.cuda_event), self._stream())
            if rc != 0:
                raise RuntimeError("B200 execution failed: " + _native.last_error(self._handle, self._lib))
            if self._use_hg:        # HG_Composite.forward: the base output goes through the highlight generator
                if self._gpu_hg_out is None:
                    self._gpu_hg_out = torch.empty((1, 3, h, w), dtype=torch.float32, device=self.device)
                rc = self._lib.hdrtv_hg(self._handle, self._gpu_out.data_ptr(), h, w, self._gpu_hg_out.data_ptr(), self._stream())
                if rc != 0:
                    raise RuntimeError("B200 execution failed: " + _native.last_error(self._handle, self._lib))
                return self._gpu_hg_out, self._gpu_agcm
        return self._gpu_out, self._gpu_agcm

    @torch.inference_mode()
    def hg_stage(self, base_out: torch.Tensor) -> torch.Tensor:
        """Extension: the HG stage alone on a (1,3,H,W) base-model output (HG_Composite_arch.py:88-107) -> new float32 tensor."""
        if not self._use_hg:
            raise RuntimeError("HG weights are not installed")
        h, w = int(base_out.shape[2]), int(base_out.shape[3])
        with torch.cuda.device(self.device):
            src = base_out.to(device=self.device, dtype=self._dtype).contiguous()
            out = torch.empty((1, 3, h, w), dtype=torch.float32, device=self.device)
            self._check(self._lib.hdrtv_hg(self._handle, src.data_ptr(), h, w, out.data_ptr(), self._stream()), "hdrtv_hg")
            src.record_stream(torch.cuda.current_stream(self.device))
        return out

    def hg_time_plan(self, base_out: torch.Tensor):
        """[(launch name, ms)] of one HG pass (FP16 contexts), CUDA events between launches."""
        h, w = int(base_out.shape[2]), int(base_out.shape[3])
        with torch.cuda.device(self.device):
            src = base_out.to(device=self.device, dtype=self._dtype).contiguous()
            out = torch.empty((1, 3, h, w), dtype=torch.float32, device=self.device)
            ms = (C.c_float * 64)()
            names = C.create_string_buffer(8192)
            n = self._lib.hdrtv_hg_time_plan(self._handle, src.data_ptr(), h, w, out.data_ptr(), ms, 64, names, 8192, self._stream())
            if n < 0:
                raise RuntimeError("hdrtv_hg_time_plan failed: " + _native.last_error(self._handle, self._lib))
        return list(zip(names.value.decode().strip().split("\n"), [float(ms[i]) for i in range(n)]))

    # ------------------------------------------------------------------ letterbox (gui_scaling.py:228-244)
    @torch.inference_mode()
    def letterbox_bgr(self, frame, out_w: int, out_h: int):
        """`_letterbox_bgr(frame, out_w, out_h)` on the GPU: aspect-preserving resize (INTER_AREA when shrinking,
        INTER_CUBIC when enlarging) centred on a black canvas.  uint8 HxWx3 numpy array in -> numpy array out (like the
        reference's host function), or CUDA uint8 tensor in -> CUDA tensor out (feed it to preprocess_device /
        process_rgb48 without leaving the device).  A frame that already has the requested size is returned as is."""
        is_np = isinstance(frame, np.ndarray)
        if is_np:
            if frame.dtype != np.uint8 or frame.ndim != 3 or frame.shape[2] != 3:
                raise ValueError("frame must be a uint8 HxWx3 BGR array")
        elif not (isinstance(frame, torch.Tensor) and frame.device.type == "cuda" and frame.dtype == torch.uint8
                  and frame.dim() == 3 and frame.shape[2] == 3):
            raise ValueError("frame must be a uint8 HxWx3 numpy array or CUDA tensor")
        h, w = int(frame.shape[0]), int(frame.shape[1])
        out_w, out_h = int(out_w), int(out_h)
        if w == out_w and h == out_h:
            return frame
        with torch.cuda.device(self.device):
            src = torch.from_numpy(np.ascontiguousarray(frame)).to(self.device, non_blocking=False) if is_np else frame.contiguous()
            dst = torch.empty((out_h, out_w, 3), dtype=torch.uint8, device=self.device)
            self._check(self._lib.hdrtv_letterbox_bgr(self._handle, src.data_ptr(), h, w, dst.data_ptr(), out_h, out_w, self._stream()),
                        "hdrtv_letterbox_bgr")
            if not is_np:
                src.record_stream(torch.cuda.current_stream(self.device))
                return dst
            return dst.cpu().numpy()

    # ------------------------------------------------------------------ postprocess (hdrtvnet_torch.py:2352-2368)
    @torch.inference_mode()
    def postprocess(self, output):
        if isinstance(output, (tuple, list)):
            output = output[0]
        if output.dtype not in (torch.float16, torch.float32):
            output = output.float()
        h, w = int(output.shape[-2]), int(output.shape[-1])
        with torch.cuda.device(self.device):
            self._ensure_buffers(h, w) if self._buf_hw != (h, w) else None
            src = output.contiguous()
            dt = _native.FP16 if src.dtype == torch.float16 else _native.FP32
            self._check(self._lib.hdrtv_pack_bgr24(self._handle, src.data_ptr(), dt, h, w, self._gpu_u8.data_ptr(),
                                                   self._stream()), "hdrtv_pack_bgr24")
            self._pin_output.copy_(self._gpu_u8, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        return self._pin_output.numpy()

    # ------------------------------------------------------------------ public API (hdrtvnet_torch.py:2373-2395)
    @torch.inference_mode()
    def process(self, frame_bgr):
        tensor, cond = self.preprocess(frame_bgr)
        out = self.infer((tensor, cond))
        return self.postprocess(out)

    @torch.inference_mode()
    def process_timed(self, frame_bgr):
        t0 = time.perf_counter()
        tensor, cond = self.preprocess(frame_bgr)
        torch.cuda.synchronize(self.device)
        t1 = time.perf_counter()
        out = self.infer((tensor, cond))
        torch.cuda.synchronize(self.device)
        t2 = time.perf_counter()
        output = self.postprocess(out)
        t3 = time.perf_counter()
        return output, (t1 - t0) * 1000.0, (t2 - t1) * 1000.0, (t3 - t2) * 1000.0

    # ------------------------------------------------------------------ one-call frame path (extension)
    @torch.inference_mode()
    def process_rgb48(self, frame_bgr, serial: bool = False, transfer: str = "identity", input_ready: bool = False,
                      checksum: bool = False):
        """BGR24 frame -> RGB48 frame in a pinned ring slot through ONE C-ABI call (``hdrtv_process``): what the
        playback / export loops do per frame with ``preprocess`` -> ``infer`` -> ``_tensor_to_rgb48_bytes``
        (gui_pipeline_worker_frame_processing.py:168-331, gui_pipeline_worker_feeders.py:193-249), bit-identical to
        that sequence.  ``frame_bgr``: uint8 HxWx3 numpy array (pinned memory is DMA-copied straight from the array,
        which must stay untouched until the returned frame is ready) or a CUDA uint8 tensor.  Returns a
        ``PinnedFrame`` (wait_ready / buffer_view / release).  ``serial=True`` keeps every stage on the current
        stream (lowest single-frame latency); the default overlaps frame k+1's copy-in / preprocess / classifier and
        frame k-1's copy-out with frame k's network.  ``checksum=True``: the pack kernel also computes the frame's descriptor
        checksum (``sharding.frame_checksum``) on the device; read it with ``PinnedFrame.checksum()``."""
        from .feeders import PinnedFrame, PinnedRing, pq_code_table
        if isinstance(frame_bgr, torch.Tensor):
            if frame_bgr.device.type != "cuda" or frame_bgr.dtype != torch.uint8 or frame_bgr.dim() != 3 or frame_bgr.shape[2] != 3:
                raise ValueError("frame_bgr must be a CUDA uint8 HxWx3 tensor or a uint8 HxWx3 numpy array")
            src = frame_bgr.contiguous()
            ptr = src.data_ptr()
        else:
            if not isinstance(frame_bgr, np.ndarray) or frame_bgr.dtype != np.uint8 or frame_bgr.ndim != 3 or frame_bgr.shape[2] != 3:
                raise ValueError("frame_bgr must be a uint8 HxWx3 BGR array")
            src = np.ascontiguousarray(frame_bgr)
            ptr = src.ctypes.data
        if transfer not in ("identity", "pq1000"):
            raise ValueError("transfer must be 'identity' or 'pq1000'")
        h, w = int(src.shape[0]), int(src.shape[1])
        with torch.cuda.device(self.device):
            self._ensure_buffers(h, w)
            if transfer == "pq1000" and not self._lut_set:
                lut = np.ascontiguousarray(pq_code_table(1000.0))
                self._check(self._lib.hdrtv_set_transfer_lut(self._handle, lut.ctypes.data, lut.size), "hdrtv_set_transfer_lut")
                self._lut_set = True
            if self._rgb48_ring is None:
                self._rgb48_ring = PinnedRing(self.device, self._rgb48_ring_frames)
            slot = self._rgb48_ring.acquire((h, w, 3))
            slot["source"] = src                       # keeps the input frame alive until the slot is reused
            mode = (_native.COND_ZERO if self._fast_zero_condition else
                    (_native.COND_BILINEAR if self._fast_condition_resize else _native.COND_BICUBIC_AA))
            flags = (_native.PROCESS_SERIAL if serial else 0) | (_native.PROCESS_INPUT_READY if input_ready else 0)
            if not self._proc_dirty:                   # first one-call frame after preprocess()/infer() used the context
                flags |= _native.PROCESS_RESYNC
                if self._side is not None:
                    torch.cuda.current_stream(self.device).wait_stream(self._side)
            self._proc_dirty = True
            tr = _native.TRANSFER_LUT if transfer == "pq1000" else _native.TRANSFER_IDENTITY
            self._cls_ready = False
            slot["has_checksum"] = bool(checksum)
            rc = self._lib.hdrtv_process_ex(self._handle, ptr, h, w, slot["tensor"].data_ptr(), mode, tr, flags,
                                            C.c_void_p(slot["checksum"].data_ptr()) if checksum else None,
                                            C.c_void_p(slot["event"].cuda_event), self._stream())
            if rc != 0:
                slot["free"].set()
                raise RuntimeError("B200 execution failed: " + _native.last_error(self._handle, self._lib))
        return PinnedFrame(slot, slot["event"])

    def set_rgb48_ring_frames(self, n: int):
        """Depth of process_rgb48's pinned RGB48 ring (frames that may be in flight + 1); takes effect when the ring is
        (re)created.  The reference's ring holds 3 frames (gui_pipeline_worker_feeders.py:28-36)."""
        n = max(2, min(16, int(n)))
        if self._rgb48_ring is not None and self._rgb48_ring.ring_frames != n:
            torch.cuda.synchronize(self.device)
            self._rgb48_ring = None
        self._rgb48_ring_frames = n

    def _warmup(self):
        h, w = self.expected_hw or (1080, 1920)
        dummy = np.zeros((h, w, 3), dtype=np.uint8)
        t0 = time.perf_counter()
        for _ in range(int(self._warmup_passes)):
            self.infer(self.preprocess(dummy))
        torch.cuda.synchronize(self.device)
        dt = time.perf_counter() - t0
        print(f"  Warmup done: {self._warmup_passes} passes in {dt:.2f}s")

    @torch.inference_mode()
    def warmup_compile(self, width=1920, height=1080):
        if not self._compiled:     # hdrtvnet_torch.py:2413-2414 — nothing is JIT-compiled here
            return

    def end_profiling(self):
        return None

    # ------------------------------------------------------------------ introspection (tests / bench)
    def launch_count(self) -> int:
        return int(self._lib.hdrtv_launch_count(self._handle))

    def workspace_bytes(self) -> int:
        return int(self._lib.hdrtv_workspace_bytes(self._handle))

    def debug_tensors(self) -> dict:
        """Named intermediates of the last infer() as (C,H,W) fp32 numpy arrays (synchronises)."""
        self._need_debug_library()
        out = {}
        n = self._lib.hdrtv_debug_tensor_count(self._handle)
        name = C.create_string_buffer(128)
        c, h, w = C.c_int(), C.c_int(), C.c_int()
        for i in range(n):
            self._lib.hdrtv_debug_tensor_info(self._handle, i, name, 128, C.byref(c), C.byref(h), C.byref(w))
            arr = np.empty((c.value, h.value, w.value), dtype=np.float32)
            self._check(self._lib.hdrtv_debug_tensor_read(self._handle, i, arr.ctypes.data), "debug read")
            out[name.value.decode()] = arr
        return out

    def time_plan(self, input_cond):
        """[(launch name, ms)] of one fp16 infer, CUDA events between launches."""
        tensor, cond = input_cond
        h, w = int(tensor.shape[2]), int(tensor.shape[3])
        self._ensure_buffers(h, w)
        ms = (C.c_float * 256)()
        names = C.create_string_buffer(32768)
        n = self._lib.hdrtv_time_plan(self._handle, tensor.data_ptr(), cond.data_ptr(), h, w, self._gpu_out.data_ptr(),
                                      self._gpu_agcm.data_ptr(), ms, 256, names, 32768, self._stream())
        if n < 0:
            raise RuntimeError("hdrtv_time_plan failed: " + _native.last_error(self._handle, self._lib))
        return list(zip(names.value.decode().strip().split("\n"), [float(ms[i]) for i in range(n)]))

    def mma_probe(self, n, layout=0, vary=1, iters=2000, blocks=1, nacc=1):
        self._need_debug_library()
        cyc = C.c_float()
        self._check(self._lib.hdrtv_mma_probe(self._handle, n, layout, vary, iters, blocks, nacc, C.byref(cyc)), "hdrtv_mma_probe")
        return float(cyc.value)

    def probe(self, kind, n=64, iters=2000, blocks=1, nwarps=4, nmma=4, groups=1, trace=False):
        self._need_debug_library()
        cyc = C.c_float()
        tr = np.zeros(256, dtype=np.int64) if trace else None
        self._check(self._lib.hdrtv_probe(self._handle, kind, n, iters, blocks, nwarps, nmma, groups, C.byref(cyc),
                                          tr.ctypes.data if trace else None), "hdrtv_probe")
        if trace:
            return float(cyc.value), tr.reshape(16, 4, 4)
        return float(cyc.value)

    def debug_layer(self, layer: str, x: np.ndarray, stride: int = 1) -> np.ndarray:
        """One named conv / linear layer of the FP32 (and INT8 fake-quant) path on host data: (Cin,H,W) -> (Cout,Ho,Wo)."""
        self._need_debug_library()
        x = np.ascontiguousarray(x, dtype=np.float32)
        cin, h, w = x.shape
        wt = self._layer_shapes[layer]
        ks = wt[2] if len(wt) == 4 else 1
        ho, wo = (h + 2 * (ks // 2) - ks) // stride + 1, (w + 2 * (ks // 2) - ks) // stride + 1
        out = np.empty((wt[0], ho, wo), dtype=np.float32)
        self._check(self._lib.hdrtv_debug_layer(self._handle, layer.encode(), x.ctypes.data, cin, h, w, stride, out.ctypes.data),
                    "hdrtv_debug_layer")
        return out

    def debug_conv_i8(self, layer: str, q: np.ndarray, stride: int = 1):
        """One W8A8 layer through its kind::i8 launch on uint8 codes (C,H,W): (raw int32 accumulators (N,Ho,Wo), float output)."""
        self._need_debug_library()
        q = np.ascontiguousarray(q, dtype=np.uint8)
        cin, h, w = q.shape
        n = self._layer_shapes[layer][0]
        ho, wo = ((h - 1) // 2 + 1, (w - 1) // 2 + 1) if stride == 2 else (h, w)
        acc = np.empty((n, ho, wo), dtype=np.int32)
        out = np.empty((n // 4, 2 * ho, 2 * wo) if n == 128 else (n, ho, wo), dtype=np.float32)
        self._check(self._lib.hdrtv_debug_conv_i8(self._handle, layer.encode(), q.ctypes.data, cin, h, w, stride, acc.ctypes.data,
                                                  out.ctypes.data), "hdrtv_debug_conv_i8")
        return acc, out

    def chain_trace(self, agcm=False, index=0):
        self._need_debug_library()
        tr = np.zeros(64 * 8 * 8, dtype=np.int64)
        self._check(self._lib.hdrtv_chain_trace(self._handle, 1 if agcm else 0, index, tr.ctypes.data), "hdrtv_chain_trace")
        return tr.reshape(64, 8, 8)

    def conv_selftest(self, kind, cin, cout, h, w, flags=0):
        self._need_debug_library()
        mx, ref = C.c_float(), C.c_float()
        self._check(self._lib.hdrtv_conv_selftest(self._handle, kind, cin, cout, h, w, flags, C.byref(mx), C.byref(ref)),
                    "hdrtv_conv_selftest")
        return float(mx.value), float(ref.value)

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self._lib.hdrtv_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
