"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the header declares, host logic of the
drop-in class, frame sharding incl. a world_size-2 gloo run, synthetic-frame determinism."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import REPO

import hdr_realtime_video_pipeline_b200 as hb
from hdr_realtime_video_pipeline_b200 import _native, sharding
from hdr_realtime_video_pipeline_b200 import build as hbuild


@pytest.fixture(scope="session", autouse=True)
def built_library():
    hbuild.build()


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(REPO, "include", "hdrtv_b200.h")).read()
    declared = set(re.findall(r"\b(hdrtv_[a-z0-9_]+)\s*\(", header))
    declared -= {"hdrtv_t"}
    assert len(declared) >= 15
    lib = _native.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/hdrtv_b200.h but not exported"
    assert declared == set(_native.EXPORTED_SYMBOLS)
    assert b"sm_100a" in lib.hdrtv_version()
    # the product ABI is the path's entry points only: no debug / probe / self-test symbol is exported by the product library
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (hdrtv_[a-z0-9_]+)", out))
    assert exported == declared, exported ^ declared


def test_test_library_exports_the_debug_entry_points():
    """include/hdrtv_b200_test.h: debug / self-test / probe entry points live in libhdrtv_b200_test.so (same sources,
    -DHDRTV_TEST_EXPORTS), next to the whole product ABI."""
    header = open(os.path.join(REPO, "include", "hdrtv_b200_test.h")).read()
    declared = set(re.findall(r"\b(hdrtv_[a-z0-9_]+)\s*\(", header)) - {"hdrtv_t"}
    assert declared == set(_native.TEST_EXPORTED_SYMBOLS)
    lib = _native.load_test()
    for name in sorted(declared | set(_native.EXPORTED_SYMBOLS)):
        assert hasattr(lib, name), name
    assert b"test build" in lib.hdrtv_version()


def test_header_constants_match_the_binding():
    """The enum values of include/hdrtv_b200.h are restated in _native.py (ctypes has no header parser): keep them equal."""
    header = open(os.path.join(REPO, "include", "hdrtv_b200.h")).read()
    consts = {k: int(v) for k, v in re.findall(r"\b(HDRTV_[A-Z0-9_]+)\s*=\s*(\d+)", header)}
    want = {"HDRTV_FP32": _native.FP32, "HDRTV_FP16": _native.FP16, "HDRTV_COND_BICUBIC_AA": _native.COND_BICUBIC_AA,
            "HDRTV_COND_ZERO": _native.COND_ZERO, "HDRTV_COND_BILINEAR": _native.COND_BILINEAR,
            "HDRTV_TRANSFER_IDENTITY": _native.TRANSFER_IDENTITY, "HDRTV_TRANSFER_LUT": _native.TRANSFER_LUT,
            "HDRTV_PROCESS_SERIAL": _native.PROCESS_SERIAL, "HDRTV_PROCESS_INPUT_READY": _native.PROCESS_INPUT_READY,
            "HDRTV_PROCESS_RESYNC": _native.PROCESS_RESYNC}
    for k, v in want.items():
        assert consts.get(k) == v, k
    # hdrtv_process is the one-call entry of SURVEY §8b: BGR24 in, RGB48 out
    lib = _native.load()
    assert lib.hdrtv_process.argtypes is not None and len(lib.hdrtv_process.argtypes) == 10


def test_one_call_path_needs_cuda_and_valid_frames():
    """process_rgb48 has no CPU path either: constructing the backend without CUDA raises (hdrtvnet_torch.py:1678-1690)."""
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError):
        hb.HDRTVNetB200(os.path.join(REPO, "tests", "golden", "weights_hr.npz"), device="cuda", precision="fp16")


def test_library_contains_blackwell_sass():
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, cp.async.bulk -> UBLKCP (B200_PROFILING.md)."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _native.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UBLKCP"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16" not in sass          # no legacy mma.sync path


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    w = os.path.join(REPO, "tests", "golden", "weights_hr.npz")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hb.HDRTVNetB200(w, device="auto", warmup_passes=0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hb.HDRTVNetB200(w, device="cpu", warmup_passes=0)
    with pytest.raises(RuntimeError):
        hb.RGB48Packer()


def test_argument_validation_matches_reference():
    w = os.path.join(REPO, "tests", "golden", "weights_hr.npz")
    with pytest.raises(ValueError, match="device must be one of"):          # hdrtvnet_torch.py:1690
        hb.HDRTVNetB200(w, device="tpu", warmup_passes=0)
    with pytest.raises(ValueError, match="precision must be one of"):       # hdrtvnet_torch.py:1694-1695
        hb.HDRTVNetB200(w, precision="bf16", warmup_passes=0)


def test_state_dict_loading_variants(tmp_path, weights_hr):
    state, arch = hb.load_state_dict_any(os.path.join(REPO, "tests", "golden", "weights_hr.npz"))
    assert len(state) == 264 and arch == {}
    assert sum(v.size for v in state.values()) == 591158                     # 591 158 parameters
    # torch checkpoint, module.-prefixed, wrapped as a source checkpoint with architecture metadata
    wrapped = {"state_dict": {"module." + k: torch.from_numpy(v) for k, v in weights_hr.items()},
               "architecture": {"classifier": "color_condition", "le_arch": None}}
    p = tmp_path / "ckpt.pt"
    torch.save(wrapped, p)
    state2, arch2 = hb.load_state_dict_any(str(p))
    assert set(state2) == set(state) and arch2["classifier"] == "color_condition"
    assert np.array_equal(state2["LE.HR_conv1.weight"], state["LE.HR_conv1.weight"])
    with pytest.raises(FileNotFoundError):
        hb.load_state_dict_any(str(tmp_path / "missing.pt"))


def test_frame_chunks_tile_the_clip():
    for n, g in ((2400, 8), (2400, 4), (2400, 2), (17, 4), (3, 8), (0, 2)):
        chunks = [hb.frame_chunk(n, r, g) for r in range(g)]
        assert chunks[0][0] == 0 and chunks[-1][1] == n
        assert all(chunks[i][1] == chunks[i + 1][0] for i in range(g - 1))
        sizes = [b - a for a, b in chunks]
        assert max(sizes) - min(sizes) <= 1
    assert hb.frame_chunk(2400, 3, 8) == (900, 1200)
    with pytest.raises(ValueError):
        hb.frame_chunk(10, 2, 2)


def test_descriptor_merge_and_checksum():
    a = np.arange(24, dtype=np.uint16).reshape(2, 4, 3)
    b = a.copy()
    b[0, 0, 0], b[0, 0, 1] = b[0, 0, 1], b[0, 0, 0]
    assert sharding.frame_checksum(a) != sharding.frame_checksum(b)          # order-sensitive
    recs = [{"first_frame": 2, "n_frames": 2, "descriptors": [(2, 7), (3, 8)]},
            {"first_frame": 0, "n_frames": 2, "descriptors": [(0, 5), (1, 6)]}]
    assert sharding.merge_descriptors(recs) == [(0, 5), (1, 6), (2, 7), (3, 8)]
    with pytest.raises(ValueError):
        sharding.merge_descriptors([recs[0], {"first_frame": 5, "n_frames": 1, "descriptors": [(5, 1)]}])


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {repo!r})
import torch.distributed as dist
from hdr_realtime_video_pipeline_b200 import sharding
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
first, last = sharding.frame_chunk(11, rank, world)
rec = dict(rank=rank, first_frame=first, n_frames=last - first, elapsed_s=0.1 * (rank + 1),
           descriptors=[(i, i * i) for i in range(first, last)])
recs = sharding.gather_run_records(rec)
merged = sharding.merge_descriptors(recs)
assert [d[0] for d in merged] == list(range(11)), merged
assert max(r["elapsed_s"] for r in recs) == 0.1 * world
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_gloo_world_size_2_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(repo=REPO))
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out
        assert "ok" in out


class _FakePayload:
    def __init__(self, arr):
        self._a = np.ascontiguousarray(arr)
        self.released = False

    def wait_ready(self):
        pass

    def buffer_view(self):
        return memoryview(self._a).cast("B")

    def release(self):
        self.released = True


class _FakeProcessor:
    """preprocess/infer stand-in: the 'network output' of frame i is a uint16 ramp offset by the frame's first byte."""
    def preprocess(self, frame):
        return frame

    def infer(self, frame):
        return frame


def _fake_pack(frame):
    h, w = frame.shape[:2]
    base = (np.arange(h * w * 3, dtype=np.uint32) % 65000).astype(np.uint16).reshape(h, w, 3)
    return _FakePayload(base + np.uint16(frame[0, 0, 0]))


def _fake_frames(i):
    return np.full((6, 10, 3), i, dtype=np.uint8)


def test_export_writer_orders_frames_and_ffmpeg_contract(tmp_path):
    out = tmp_path / "clip.rgb48"
    rec = hb.export_clip(_FakeProcessor(), _fake_frames, 7, str(out), pack=_fake_pack)
    assert rec["n_frames"] == 7 and [d[0] for d in rec["descriptors"]] == list(range(7))
    data = np.fromfile(out, dtype=np.uint16).reshape(7, 6, 10, 3)
    for i in range(7):
        assert np.array_equal(data[i], _fake_pack(_fake_frames(i))._a)
    assert sharding.merge_descriptors([rec])[3][1] == sharding.frame_checksum(data[3])
    args = hb.ffmpeg_rawvideo_args(3840, 2160, 23.976, "out.mov")
    joined = " ".join(args)                                                  # src/gui_export.py:966-1023
    for token in ("-f rawvideo", "-pix_fmt rgb48le", "-s:v 3840x2160", "-color_range pc", "-colorspace bt2020nc",
                  "-color_trc smpte2084", "-color_primaries bt2020", "-c:v prores_ks", "-profile:v 3",
                  "transferin=smpte2084", "format=yuv422p10le"):
        assert token in joined, token
    with pytest.raises(ValueError):
        hb.Rgb48RawWriter(str(out), 7, 6, 10, create=False).write(9, b"")


_EXPORT_WORKER = r"""
import os, sys
sys.path.insert(0, {repo!r})
sys.path.insert(0, os.path.join({repo!r}, "tests"))
import torch.distributed as dist
import hdr_realtime_video_pipeline_b200 as hb
from hdr_realtime_video_pipeline_b200 import sharding
from test_host_cpu import _FakeProcessor, _fake_frames, _fake_pack
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
rec = hb.export_clip(_FakeProcessor(), _fake_frames, 9, {out!r}, rank=rank, world_size=world, pack=_fake_pack,
                     barrier=dist.barrier)
recs = sharding.gather_run_records(rec)
merged = sharding.merge_descriptors(recs)
assert [d[0] for d in merged] == list(range(9)), merged
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_gloo_world_size_2_sharded_export(tmp_path):
    out = tmp_path / "clip2.rgb48"
    script = tmp_path / "export_worker.py"
    script.write_text(_EXPORT_WORKER.format(repo=REPO, out=str(out)))
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29534")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        o, _ = p.communicate(timeout=180)
        assert p.returncode == 0, o
    data = np.fromfile(out, dtype=np.uint16).reshape(9, 6, 10, 3)
    for i in range(9):
        assert np.array_equal(data[i], _fake_pack(_fake_frames(i))._a)      # both ranks' chunks, in clip order


def test_synth_frames_deterministic_and_typed():
    for i in range(4):
        a, b = hb.synth_frame(i, 36, 52), hb.synth_frame(i, 36, 52)
        assert a.dtype == np.uint8 and a.shape == (36, 52, 3) and a.flags["C_CONTIGUOUS"]
        assert np.array_equal(a, b)
    assert hb.synth_frame(2, 8, 8).max() == 0                    # class C: the reference's zero frame
    assert (hb.synth_frame(3, 64, 64) == 255).mean() > 0.9       # class D: near-white
    assert [i for i, _ in hb.synth_clip(3, 8, 8, first_frame=5)] == [5, 6, 7]


def test_pq_code_table_matches_oracle_formula():
    from oracle import hdrtvnet_oracle as O
    lut = hb.pq_code_table(1000.0)
    assert lut.shape == (0x3C01,) and lut[0] == 0
    x = np.arange(0x3C01, dtype=np.uint16).view(np.float16).astype(np.float32)
    ref = O.pack_rgb48_pq(np.stack([x, x, x])[None, :, None, :], 1000.0)[0, :, 0]
    assert np.array_equal(lut, ref)
    assert abs(int(lut[-1]) - round(0.7518 * 65535)) < 40        # PQ(1000 nit) ~ 0.752
