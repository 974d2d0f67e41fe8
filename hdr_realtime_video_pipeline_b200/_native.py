"""ctypes binding of the C ABI in include/hdrtv_b200.h (libhdrtv_b200.so, built in-tree by build.py).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhdrtv_b200.so")
TEST_LIB_PATH = os.path.join(_HERE, "libhdrtv_b200_test.so")      # product ABI + debug / probe entry points (tests, scripts)

FP32, FP16 = 0, 1
COND_BICUBIC_AA, COND_ZERO, COND_BILINEAR = 0, 1, 2
TRANSFER_IDENTITY, TRANSFER_LUT = 0, 1
PROCESS_SERIAL, PROCESS_INPUT_READY, PROCESS_RESYNC = 1, 2, 4


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("precision", C.c_int)]


class TensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.POINTER(C.c_float)), ("ndim", C.c_int), ("shape", C.c_int64 * 4)]


_SIGNATURES = {
    "hdrtv_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "hdrtv_destroy": (None, [C.c_void_p]),
    "hdrtv_set_weights": (C.c_int, [C.c_void_p, C.POINTER(TensorDesc), C.c_int]),
    "hdrtv_set_act_quant": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                      C.POINTER(C.c_int), C.c_int]),
    "hdrtv_prepare": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "hdrtv_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "hdrtv_preprocess": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "hdrtv_preprocess_classify": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "hdrtv_infer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hdrtv_classify": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "hdrtv_infer_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                 C.c_void_p, C.c_void_p]),
    "hdrtv_pack_rgb48": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "hdrtv_process": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                C.c_void_p]),
    "hdrtv_process_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "hdrtv_process_flush": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hdrtv_process_output": (C.c_void_p, [C.c_void_p, C.c_int]),
    "hdrtv_set_transfer_lut": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "hdrtv_letterbox_bgr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "hdrtv_set_hg_weights": (C.c_int, [C.c_void_p, C.POINTER(TensorDesc), C.c_int]),
    "hdrtv_hg": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "hdrtv_hg_time_plan": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_float), C.c_int,
                                     C.c_char_p, C.c_int, C.c_void_p]),
    "hdrtv_pack_bgr24": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "hdrtv_last_error": (C.c_char_p, [C.c_void_p]),
    "hdrtv_launch_count": (C.c_long, [C.c_void_p]),
    "hdrtv_time_plan": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.POINTER(C.c_float), C.c_int, C.c_char_p, C.c_int, C.c_void_p]),
    "hdrtv_version": (C.c_char_p, []),
}
# include/hdrtv_b200_test.h: exported by libhdrtv_b200_test.so only
_TEST_SIGNATURES = {
    "hdrtv_debug_layer": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "hdrtv_debug_conv_i8": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "hdrtv_debug_tensor_count": (C.c_int, [C.c_void_p]),
    "hdrtv_debug_tensor_info": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "hdrtv_debug_tensor_read": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "hdrtv_conv_selftest": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "hdrtv_mma_probe": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "hdrtv_probe": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_void_p]),
    "hdrtv_chain_trace": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
TEST_EXPORTED_SYMBOLS = tuple(_TEST_SIGNATURES)
_lib = None
_test_lib = None


def _open(path, signatures):
    if not os.path.isfile(path):
        raise RuntimeError(
            f"{path} not found: the CUDA extension is required (no CPU fallback). "
            "Build it with `python -m hdr_realtime_video_pipeline_b200.build` or __graft_entry__.build().")
    lib = C.CDLL(path)
    for name, (res, args) in signatures.items():
        fn = getattr(lib, name)   # AttributeError here means the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    return lib


def load():
    """Load the product library libhdrtv_b200.so and attach signatures.  Raises RuntimeError when it is absent."""
    global _lib
    if _lib is None:
        _lib = _open(LIB_PATH, _SIGNATURES)
    return _lib


def load_test():
    """Load the test build (product ABI + debug / probe entry points).  Contexts are tied to the library that made them."""
    global _test_lib
    if _test_lib is None:
        _test_lib = _open(TEST_LIB_PATH, {**_SIGNATURES, **_TEST_SIGNATURES})
    return _test_lib


def last_error(handle, lib=None) -> str:
    msg = (lib or load()).hdrtv_last_error(handle)
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, handle, what: str, lib=None):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {last_error(handle, lib)}")
