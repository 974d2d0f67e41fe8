"""GPU tests against the UNMODIFIED reference running on the same box (``baseline/_ref``, installed by
scripts/install_reference.py; git-ignored, it travels with the snapshot).  They are skipped when that copy is absent;
the committed fixtures (tests/golden, test_gpu_parity.py) cover the same ground from CPU runs of the reference.

  * FP16 parity at the BASELINE config sizes against the reference's CUDA FP16 eager path
    (``HDRTVNetTorch(device="cuda", precision="fp16")``: channels_last + cudnn.benchmark, hdrtvnet_torch.py:1573-1592,
    2164-2167) over the FULL frame: the tolerance BASELINE.json states (2e-3 max-abs) is defined against exactly this.
  * P7: the reference's own ``PipelineWorkerFrameProcessingMixin._process_frame`` and ``PipelineWorkerFeedersMixin.
    _hdr_feeder_fn`` (gui_pipeline_worker_frame_processing.py:168-331, gui_pipeline_worker_feeders.py:313-496) run
    UNCHANGED on top of ``HDRTVNetB200`` / ``tensor_to_rgb48_bytes``.
"""
import os
import queue
import threading

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

import hdr_realtime_video_pipeline_b200 as hb  # noqa: E402
from oracle import hdrtvnet_oracle as O  # noqa: E402
from oracle import reference_loader as RL  # noqa: E402
from test_gpu_parity import FP16_TOL, FP32_TOL, fp16_gate  # noqa: E402

W_HR = os.path.join(GOLDEN, "weights_hr.npz")
REF = RL.load()
needs_ref = pytest.mark.skipif(REF is None, reason="baseline/_ref not installed (scripts/install_reference.py)")


@pytest.fixture(scope="module")
def rand0_ckpt(tmp_path_factory, weights_rand0):
    path = tmp_path_factory.mktemp("w") / "rand0.pt"
    torch.save({k: torch.from_numpy(np.asarray(v)) for k, v in weights_rand0.items()}, path)
    return str(path)


@pytest.fixture(scope="module")
def ref_nets(rand0_ckpt):
    made = {}

    def get(wname, precision):
        key = (wname, precision)
        if key not in made:
            path = REF.weights("HR.pt") if wname == "hr" else rand0_ckpt
            made[key] = REF.HDRTVNetTorch(path, device="cuda", precision=precision, compile_model=False, use_hg=False,
                                          warmup_passes=0)
        return made[key]

    yield get
    made.clear()
    torch.cuda.empty_cache()


@pytest.fixture(scope="module")
def our_nets(weights_rand0):
    made = {}

    def get(wname, precision):
        key = (wname, precision)
        if key not in made:
            made[key] = hb.HDRTVNetB200(W_HR if wname == "hr" else weights_rand0, device="cuda", precision=precision,
                                        warmup_passes=0, use_hg=False)
        return made[key]

    yield get
    for n in made.values():
        n.close()


def _ref_run(net, frame):
    with torch.inference_mode():
        t, c = net.preprocess(frame)
        res = net.infer((t.clone(), c.clone()))
        out, agcm = res[0], res[1]
        torch.cuda.synchronize()
        return out.float().contiguous().cpu().numpy(), agcm.float().contiguous().cpu().numpy(), t.float().contiguous().cpu().numpy()


LIVE_CASES = [("hr", "noise", 0, 540, 960), ("hr", "ramps", 1, 540, 960), ("rand0", "noise", 0, 540, 960),
              ("hr", "noise", 0, 1080, 1920), ("hr", "ramps", 1, 1080, 1920), ("hr", "white_salt", 3, 1080, 1920),
              ("hr", "black", 2, 1080, 1920), ("rand0", "noise", 0, 1080, 1920),
              ("hr", "noise", 0, 2160, 3840), ("hr", "ramps", 1, 2160, 3840), ("hr", "white_salt", 3, 2160, 3840)]


@needs_ref
@pytest.mark.parametrize("case", LIVE_CASES, ids=lambda c: "%s_%s_%d_%dx%d" % c)
def test_fp16_full_frame_parity_against_the_reference_cuda_fp16_path(ref_nets, our_nets, parity_log, case):
    wname, cls, idx, h, w = case
    frame = hb.synth_frame(idx, h, w, cls)
    ref16, ragcm16, rx16 = _ref_run(ref_nets(wname, "fp16"), frame)
    ref32, _, _ = _ref_run(ref_nets(wname, "fp32"), frame)
    net = our_nets(wname, "fp16")
    x, cond = net.preprocess(frame)
    out, agcm = net.infer((x, cond))
    torch.cuda.synchronize()
    assert np.array_equal(x.float().cpu().numpy(), rx16)                        # P1(ii) bit-exact against the reference's CUDA path
    out, agcm = out.float().cpu().numpy(), agcm.float().cpu().numpy()
    d16, d32, dref = np.abs(out - ref16).max(), np.abs(out - ref32).max(), np.abs(ref16 - ref32).max()
    a16 = np.abs(agcm - ragcm16).max()
    m16 = np.abs(out - ref16).mean()
    print(f"{case}: |ours-ref16|={d16:.2e} (mean {m16:.2e}) |ours-ref32|={d32:.2e} |ref16-ref32|={dref:.2e} agcm {a16:.2e}")
    parity_log.add(test="fp16_live_full_frame", case="%s_%s_%d_%dx%d" % case, height=h, width=w, ours_vs_ref16=d16,
                   ours_vs_ref16_mean=m16, ours_vs_ref32=d32, ref16_vs_ref32=dref, agcm_vs_ref16=a16, samples=int(out.size),
                   reference="HDRTVNetTorch(device='cuda', precision='fp16') and 'fp32', unmodified reference, same box (baseline/_ref)")
    fp16_gate(d16, d32, dref, str(case), slack=2.0 ** -12)
    # RGB48 codes through the one-call path against the reference feeder's pack of the reference FP16 output
    fr = net.process_rgb48(frame)
    codes = fr.numpy().copy()
    fr.release()
    ref_codes = O.pack_rgb48(ref16.astype(np.float16))
    assert np.abs(codes.astype(np.int32) - ref_codes.astype(np.int32)).max() <= int(max(d16, FP16_TOL) * 65535) + 2


@needs_ref
@pytest.mark.parametrize("hw", [(540, 960), (1080, 1920)])
def test_fp32_full_frame_parity_against_the_reference_cuda_fp32_path(ref_nets, our_nets, parity_log, hw):
    frame = hb.synth_frame(0, hw[0], hw[1], "noise")
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # the reference does not enable TF32 either; be explicit
    try:
        ref32, ragcm, _ = _ref_run(ref_nets("hr", "fp32"), frame)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    net = our_nets("hr", "fp32")
    out, agcm = net.infer(net.preprocess(frame))
    torch.cuda.synchronize()
    d = np.abs(out.cpu().numpy() - ref32).max()
    da = np.abs(agcm.cpu().numpy() - ragcm).max()
    parity_log.add(test="fp32_live_full_frame", case="hr_noise_%dx%d" % hw, ours_vs_ref32=d, agcm_vs_ref32=da,
                   reference="HDRTVNetTorch(device='cuda', precision='fp32'), same box")
    assert d <= FP32_TOL and da <= FP32_TOL, (d, da)


# ------------------------------------------------------------------------------------------------ P7, unmodified callers
class _Widget:
    """Stand-in for MpvHDRWidget.feed_frame (gui_mpv_widget.py:706-735): takes bytes or a pinned payload."""

    def __init__(self):
        self.frames = []

    def feed_frame(self, payload):
        if hasattr(payload, "buffer_view") and hasattr(payload, "release"):
            payload.wait_ready()
            self.frames.append(bytes(payload.buffer_view()))
            payload.release()
        else:
            self.frames.append(bytes(payload))


def _make_worker(processor):
    fp = REF.frame_processing.PipelineWorkerFrameProcessingMixin
    fd = REF.feeders.PipelineWorkerFeedersMixin

    class Worker(fp, fd):
        def _preserve_display_queue_order(self):      # file playback with HDRTVNET_VIDEO_PLAYBACK_PRESERVE_ORDER=1 (gui_config.py:390):
            return True                               # the default "latest wins" queue drops frames whenever the feeder lags

    wk = Worker()
    wk._processor = processor
    wk._sdr_visible = False
    wk._sdr_mpv_widget = None
    wk._sdr_queue = None
    wk._hdr_queue = queue.Queue(maxsize=2)
    wk._input_is_hdr = False
    wk._hdr_drop_until_frame = 0
    wk._sdr_drop_until_frame = 0
    wk._video_playback_buffer_frames = 2           # -> staging pool of 4 tensors
    wk._stop_flag = False
    wk._seek_frame = None
    wk._capture_target = None
    return wk


def _play(processor, frames, pack_fn=None):
    """The reference's playback loop body (PipelineWorker.run -> _process_frame) feeding the reference's feeder thread."""
    wk = _make_worker(processor)
    widget = _Widget()
    old = REF.feeders._tensor_to_rgb48_bytes
    if pack_fn is not None:
        REF.feeders._tensor_to_rgb48_bytes = pack_fn          # the one-line integration: INTEGRATION.md
    try:
        errors = []

        def feeder():
            try:
                wk._hdr_feeder_fn(wk._hdr_queue, widget, False, 60.0, True)
            except BaseException as exc:             # a dead feeder must not leave the producer spinning on a full queue
                errors.append(exc)
                wk._stop_flag = True

        th = threading.Thread(target=feeder, daemon=True)
        th.start()
        h, w = frames[0].shape[:2]
        lat = []
        for i, f in enumerate(frames):
            _, _, prepared, need_cpu, ms = wk._process_frame(frame=f, frame_idx=i, present_t=None, out_w=w, out_h=h, proc_w=w,
                                                             proc_h=h, lower_res_processing=False, mpv_w=widget, use_cuda=True)
            assert torch.is_tensor(prepared) and tuple(prepared.shape) == (1, 3, h, w) and need_cpu is False and ms > 0.0
            lat.append(ms)
        assert not errors, errors
        wk._hdr_queue.put(None, timeout=30)
        th.join(timeout=120)
        assert not th.is_alive() and not errors, errors
    finally:
        REF.feeders._tensor_to_rgb48_bytes = old
    return widget.frames, lat


@needs_ref
@pytest.mark.parametrize("precision", ["fp16", "fp32"])
def test_p7_reference_playback_loop_and_feeder_thread_run_unchanged_on_this_backend(our_nets, precision):
    """200 frames through the reference's _process_frame (CUDA events on the current stream around preprocess + infer,
    4-deep empty_like/copy_(non_blocking) staging pool, ready_event) and its _hdr_feeder_fn thread (ready_event.synchronize
    -> _tensor_to_rgb48_bytes -> feed_frame -> release), both imported from the reference and not edited:
    (a) with the reference's own _tensor_to_rgb48_bytes, (b) with this repo's drop-in tensor_to_rgb48_bytes.
    Every frame must be byte-equal to serial execution of the three calls and to process_rgb48."""
    net = our_nets("hr", precision)
    h, w = 136, 248
    frames = [hb.synth_frame(i, h, w) for i in range(200)]
    want = []
    for f in frames:
        out, _ = net.infer(net.preprocess(f))
        torch.cuda.synchronize()
        want.append(O.pack_rgb48(out.cpu().numpy()).tobytes())
    got_ref_pack, lat = _play(net, frames)
    got_our_pack, _ = _play(net, frames, pack_fn=hb.tensor_to_rgb48_bytes)
    assert len(got_ref_pack) == len(got_our_pack) == len(frames)
    bad_a = [i for i, (a, b) in enumerate(zip(got_ref_pack, want)) if a != b]
    bad_b = [i for i, (a, b) in enumerate(zip(got_our_pack, want)) if a != b]
    assert not bad_a and not bad_b, (bad_a[:8], bad_b[:8])
    one = []
    for f in frames[:24]:
        fr = net.process_rgb48(f)
        one.append(bytes(fr.buffer_view()))
        fr.release()
    assert one == want[:24]
    assert float(np.median(lat)) < 50.0


# ------------------------------------------------------------------------------------------------ INT8 Mixed QAT, live
@needs_ref
@pytest.mark.parametrize("hw", [(540, 960), (1080, 1920)])
def test_int8_mixed_against_the_reference_cuda_int8_path(parity_log, hw):
    """BASELINE config 5 on the shipping layout: the reference's eager INT8-mixed model on this GPU (W8A8Conv2d with
    compute_dtype fp16, hdrtvnet_torch.py:296-364, 1748-1963) against the kind::i8 tensor-core path on the same checkpoint."""
    ckpt = REF.weights("HR_original_int8_mixed_qat.pt")
    ref = REF.HDRTVNetTorch(ckpt, device="cuda", precision="int8-mixed", compile_model=False, use_hg=False, warmup_passes=0,
                            predequantize="off")
    net = hb.HDRTVNetB200(ckpt, device="cuda", precision="int8-mixed", warmup_passes=0, use_hg=False)
    assert net._int8_tensor_path
    worst = 0.0
    for cls, idx in (("noise", 0), ("ramps", 1)):
        frame = hb.synth_frame(idx, hw[0], hw[1], cls)
        r16, ragcm, _ = _ref_run(ref, frame)
        out, agcm = net.infer(net.preprocess(frame))
        torch.cuda.synchronize()
        d = np.abs(out.float().cpu().numpy() - r16)
        da = np.abs(agcm.float().cpu().numpy() - ragcm).max()
        print(f"INT8-mixed {cls} {hw}: mean {d.mean():.2e} max {d.max():.2e} agcm {da:.2e}")
        parity_log.add(test="int8_mixed_live", case="%s_%d_%dx%d" % (cls, idx, hw[0], hw[1]), mean_abs=d.mean(), max_abs=d.max(),
                       agcm_max_abs=da, reference="HDRTVNetTorch(device='cuda', precision='int8-mixed'), unmodified reference, same box")
        assert da <= FP16_TOL
        assert d.mean() <= 1.5e-3 and d.max() <= 3e-2
        worst = max(worst, float(d.max()))
    net.close()
    del ref
    torch.cuda.empty_cache()
