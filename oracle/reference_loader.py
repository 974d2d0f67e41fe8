"""Locate and import the UNMODIFIED reference (test / benchmark infrastructure only — never imported by the product).

Resolution order: ``baseline/_ref`` (the copy made by scripts/install_reference.py, travels to the GPU box), then
``$HDRTV_REFERENCE``.  ``/root/reference`` itself is only used by the fixture generators in the build container
(``allow_source_tree=True``): nothing that runs on the GPU box reads it.

    ref = load()                 # None when no reference tree is available
    ref.HDRTVNetTorch, ref.Ensemble_AGCM_LE, ref.feeders, ref.frame_processing, ref.weights("HR.pt")
"""
from __future__ import annotations

import os
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_cached = None


class Reference:
    def __init__(self, root: str):
        self.root = root
        src = os.path.join(root, "src")
        if src not in sys.path:
            sys.path.insert(0, src)
        # gui_pipeline_worker_feeders imports the PyQt6 mpv widget (absent here); stub the module, nothing else
        if "gui_mpv_widget" not in sys.modules:
            try:
                import PyQt6  # noqa: F401
            except Exception:
                stub = types.ModuleType("gui_mpv_widget")
                stub.MpvHDRWidget = type("MpvHDRWidget", (), {})
                sys.modules["gui_mpv_widget"] = stub
        from models.hdrtvnet_torch import HDRTVNetTorch
        from models.hdrtvnet_modules.Ensemble_AGCM_LE_arch import Ensemble_AGCM_LE
        self.HDRTVNetTorch = HDRTVNetTorch
        self.Ensemble_AGCM_LE = Ensemble_AGCM_LE
        self._feeders = None
        self._frame_processing = None

    @property
    def feeders(self):
        if self._feeders is None:
            import gui_pipeline_worker_feeders as m
            self._feeders = m
        return self._feeders

    @property
    def frame_processing(self):
        if self._frame_processing is None:
            import gui_pipeline_worker_frame_processing as m
            self._frame_processing = m
        return self._frame_processing

    def weights(self, name: str = "HR.pt") -> str:
        base = os.path.join(self.root, "src", "models", "weights", "original")
        for cand in (os.path.join(base, name), os.path.join(base, "pytorch_int8", "hr", name)):
            if os.path.isfile(cand):
                return cand
        raise FileNotFoundError(name)


def find_root(allow_source_tree: bool = False) -> str | None:
    cands = [os.path.join(REPO, "baseline", "_ref"), os.environ.get("HDRTV_REFERENCE")]
    if allow_source_tree:
        cands.append("/root/reference")
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "src", "models", "hdrtvnet_torch.py")):
            return c
    return None


def load(allow_source_tree: bool = False) -> Reference | None:
    global _cached
    if _cached is not None:
        return _cached
    root = find_root(allow_source_tree)
    if root is None:
        return None
    _cached = Reference(root)
    return _cached
