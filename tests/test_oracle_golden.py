"""Pin the CPU oracle against outputs of the REFERENCE ITSELF (tests/golden/, made by
scripts/make_golden.py).  The reference ships no golden vectors of its own."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import hdrtvnet_oracle as O

NET_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "net_*.npz")))


@pytest.mark.parametrize("name", ["pre_64x96.npz", "pre_72x100.npz", "pre_135x241.npz"])
def test_preprocess_matches_reference(name):
    g = load_golden(name)
    x, cond = O.preprocess(g["frame"], np.float32)
    assert np.array_equal(x, g["x"])                      # normalise is bit-exact (one fp32 multiply)
    assert np.abs(cond - g["cond"]).max() <= 2e-6          # 16-tap FIR, summation order differs
    x16, cond16 = O.preprocess(g["frame"], np.float16)
    assert np.array_equal(x16, g["x16"])
    # fp16 result may flip one half-ulp where the fp32 sums differ in the last bits
    assert np.abs(cond16.astype(np.float32) - g["cond16"].astype(np.float32)).max() <= 1e-3
    assert (cond16 != g["cond16"]).mean() < 0.01


def test_aa_bicubic_border_taps():
    g = load_golden("aa.npz")
    for k in [k for k in g if k.startswith("in_")]:
        out = O.cond_downsample(g[k][0], np.float32)
        ref = g["out_" + k[3:]][0]
        assert out.shape == ref.shape
        assert np.abs(out - ref).max() <= 2e-6, k


@pytest.mark.parametrize("name", NET_CASES)
def test_network_matches_reference(name, weights_hr, weights_rand0):
    g = load_golden(name)
    sd = weights_hr if name.startswith("net_hr_") else weights_rand0
    x, cond = O.preprocess(g["frame"], np.float32)
    sdf = {k: np.asarray(v, np.float32) for k, v in sd.items()}
    fea = O.classifier(sdf, cond[0])
    assert np.abs(fea - g["fea"]).max() <= 2e-5
    out, agcm_out = O.infer(sd, x, cond)
    assert np.abs(agcm_out - g["agcm_out"]).max() <= 2e-5
    assert np.abs(out - g["out"]).max() <= 5e-5
    # packs: bit-exact on identical float input ...
    assert np.array_equal(O.pack_rgb48(g["out"]), g["rgb48"])
    assert np.array_equal(O.postprocess_bgr24(g["out"]), g["bgr24"])
    # ... and within the float tolerance end to end (1e-4 ~ 6.6 codes of 65535)
    assert np.abs(O.pack_rgb48(out).astype(np.int32) - g["rgb48"].astype(np.int32)).max() <= 7


@pytest.mark.parametrize("wname", ["hr", "rand0"])
def test_network_540p_config1(wname, weights_hr, weights_rand0):
    from hdr_realtime_video_pipeline_b200.synth import synth_frame
    g = load_golden(f"net540_{wname}.npz")
    sd = weights_hr if wname == "hr" else weights_rand0
    frame = synth_frame(0, 540, 960, "noise")
    x, cond = O.preprocess(frame, np.float32)
    out, agcm_out = O.infer(sd, x, cond)
    assert np.abs(out[:, :, ::8, ::8] - g["out_sub"]).max() <= 5e-5
    assert np.abs(out[:, :, -3:, :] - g["out_last_rows"]).max() <= 5e-5      # centre-crop path (68 -> 135)
    assert np.abs(agcm_out[:, :, ::8, ::8] - g["agcm_sub"]).max() <= 2e-5
    assert abs(out.mean() - g["out_stats"][0]) <= 1e-5


def test_known_answers_from_survey(weights_hr):
    """SURVEY §8c KATs: reference, HR.pt, frame = default_rng(0) 540x960 noise."""
    frame = np.random.default_rng(0).integers(0, 256, (540, 960, 3), dtype=np.uint8)
    x, cond = O.preprocess(frame, np.float32)
    np.testing.assert_allclose(x[0, :, 0, 0], [0.7607844, 0.5098040, 0.3725490], atol=1e-6)
    np.testing.assert_allclose(cond[0, :, 0, 0], [0.5773944, 0.5742362, 0.5201451], atol=1e-5)
    sdf = {k: np.asarray(v, np.float32) for k, v in weights_hr.items()}
    fea = O.classifier(sdf, cond[0])
    np.testing.assert_allclose(fea, [0.033736, 0.011883, 0.039485, -0.013184, 0.011853, 0.041158], atol=1e-5)
    out, agcm_out = O.infer(weights_hr, x, cond)
    np.testing.assert_allclose(out[0, :, 0, 0], [0.567692, 0.496016, 0.398522], atol=1e-5)
    np.testing.assert_allclose(out[0, :, 539, 959], [0.320767, 0.406483, 0.226297], atol=1e-5)
    assert abs(float(agcm_out.mean()) - 0.441951) < 1e-5
    rgb48 = O.pack_rgb48(out)
    assert np.abs(rgb48[0, 0].astype(int) - np.array([37204, 32506, 26117])).max() <= 1
    assert abs(rgb48.astype(np.float64).mean() - 30551.15) < 0.5
    assert abs(O.postprocess_bgr24(out).astype(np.float64).mean() - 118.876) < 0.01


def test_pack_edge_values_bit_exact():
    g = load_golden("pack.npz")
    assert np.array_equal(O.pack_rgb48(g["in32"]), g["rgb48_32"])
    assert np.array_equal(O.pack_rgb48(g["in16"]), g["rgb48_16"])
    assert np.array_equal(O.postprocess_bgr24(g["in32"]), g["bgr24_32"])
    assert np.array_equal(O.postprocess_bgr24(g["in16"]), g["bgr24_16"])


def test_pq_transfer_option():
    g = load_golden("pq.npz")
    lin = g["linear_rgb"].transpose(2, 0, 1)[None]
    got = O.pack_rgb48_pq(lin, 1000.0)
    assert np.abs(got.astype(np.int32) - g["pq_rgb_u16"].astype(np.int32)).max() <= 1


def test_primitives_small_cases():
    x = np.arange(2 * 5 * 7, dtype=np.float32).reshape(2, 5, 7)
    p = O.avg_pool_3s2p1(x)
    assert p.shape == (2, 3, 4)
    assert np.isclose(p[0, 0, 0], (0 + 1 + 7 + 8) / 9.0)          # count_include_pad: always /9
    ps = O.pixel_shuffle2(np.arange(8 * 2 * 3, dtype=np.float32).reshape(8, 2, 3))
    assert ps.shape == (2, 4, 6) and ps[0, 0, 1] == 6.0 and ps[0, 1, 0] == 12.0 and ps[1, 0, 0] == 24.0
    a = O.align_to(np.arange(3 * 4 * 6, dtype=np.float32).reshape(3, 4, 6), 3, 5)   # crop: top = 0, left = 0
    assert a.shape == (3, 3, 5) and a[0, 0, 0] == 0.0
    b = O.align_to(np.ones((1, 2, 2), np.float32), 3, 4)                             # replicate pad
    assert b.shape == (1, 3, 4) and b.min() == 1.0


def test_torch_port_matches_reference_and_numpy_oracle(weights_hr):
    """oracle/torch_port.py is what bench.py times as the CPU baseline; pin it too."""
    import torch
    from oracle import torch_port as TP
    g = load_golden("net_hr_ramps_72x100.npz")
    sd = TP.to_torch_state(weights_hr)
    x, cond = TP.preprocess(g["frame"])
    out, agcm_out = TP.infer(sd, x, cond)
    assert np.abs(out.numpy() - g["out"]).max() <= 2e-5
    assert np.abs(agcm_out.numpy() - g["agcm_out"]).max() <= 2e-5
    assert np.abs(TP.process(sd, g["frame"]).astype(int) - g["bgr24"].astype(int)).max() <= 1
    assert np.abs(TP.process_rgb48(sd, g["frame"]).astype(int) - g["rgb48"].astype(int)).max() <= 3
    xo, co = O.preprocess(g["frame"], np.float32)
    oo, _ = O.infer(weights_hr, xo, co)
    assert np.abs(out.numpy() - oo).max() <= 5e-5
    torch.set_num_threads(torch.get_num_threads())


def test_bilinear_condition_matches_reference_fixture():
    """fast_condition_resize (hdrtvnet_torch.py:2268-2275) through the reference wrapper itself."""
    g = load_golden("pre_bilinear.npz")
    for hw in ("64x96", "73x101"):
        frame = g[f"frame_{hw}"]
        x, cond = O.preprocess(frame, np.float32, cond_mode="bilinear")
        assert cond.shape == g[f"cond_{hw}"].shape
        assert np.abs(cond - g[f"cond_{hw}"]).max() <= 1e-6
        _, cond16 = O.preprocess(frame, np.float16, cond_mode="bilinear")
        assert np.abs(cond16.astype(np.float32) - g[f"cond16_{hw}"].astype(np.float32)).max() <= 1e-3
    z = O.preprocess(g["frame_64x96"], np.float32, cond_mode="zero")[1]
    assert z.shape == (1, 3, 16, 24) and not z.any()


# ------------------------------------------------------------------------------------------- P8 INT8 Full-QAT layout
_INT8_STRIDE = {"LE.down_conv2": 2, "LE.CondNet4.4": 2}


def _int8_sd():
    return O.split_int8_state(load_golden("weights_int8_full_qat.npz"))


def test_int8_layers_match_reference_modules_exactly():
    """Each W8A8 layer (fake-quantised input, de-quantised int8 weights) on the input the reference's own module saw
    (hdrtvnet_torch.py:350-364).  End-to-end the fake-quantised network is chaotic at quantisation-step level, so this
    is where INT8 parity is exact."""
    g = load_golden("int8_layers_64x96.npz")
    sd = _int8_sd()
    assert sum(k.endswith(".x_scale") for k in sd) == 128                      # README: Full INT8 = 128 W8A8 layers
    for layer in [str(x) for x in g["layers"]]:
        x, ref = g[layer + "|in"][0], g[layer + "|out"][0]
        if x.ndim == 1:                                                      # W8A8Linear
            got = sd[layer + ".weight"] @ O.fake_quant_input(sd, layer, x) + sd[layer + ".bias"]
        else:
            w = sd[layer + ".weight"]
            got = O.conv2d(O.fake_quant_input(sd, layer, x), w, sd[layer + ".bias"], _INT8_STRIDE.get(layer, 1), w.shape[2] // 2)
            if got.shape != ref.shape:
                got = got[:, ::4, ::4]                                       # big maps are stored every 4th pixel
        assert np.abs(got - ref).max() <= 2e-6 * max(1.0, float(np.abs(ref).max())), layer     # fp32 summation order only


@pytest.mark.parametrize("name", ["int8_noise_64x96", "int8_ramps_72x100", "int8_white_salt_72x100"])
def test_int8_network_statistical_parity(name):
    g = load_golden(name + ".npz")
    sd = _int8_sd()
    x, c = O.preprocess(g["frame"], np.float32)
    out, agcm = O.infer(sd, x, c)
    assert np.abs(agcm - g["agcm_out"]).max() <= 5e-3                          # at most an isolated flipped bucket
    assert np.abs(agcm - g["agcm_out"]).mean() <= 1e-5
    d = np.abs(out - g["out"])
    assert d.mean() <= 8e-3 and d.max() <= 8e-2, (d.mean(), d.max())           # a few activation-quantiser steps (~0.008)


@pytest.mark.parametrize("name", ["int8mixed_noise_64x96", "int8mixed_ramps_72x100"])
def test_int8_mixed_network_statistical_parity(name):
    """The reference's shipping INT8 layout (INT8 Mixed QAT: 29 W8A8 / 78 W8A16 / 21 FP16 layers): the oracle on the raw
    checkpoint arrays against the reference's own eager run (scripts/make_golden_int8_mixed.py)."""
    g = load_golden(name + ".npz")
    sd = O.split_int8_state(load_golden("weights_int8_mixed_qat.npz"))
    assert sum(1 for k in sd if k.endswith(".x_scale")) == 29
    x, cond = O.preprocess(g["frame"], np.float32)
    out, agcm = O.infer(sd, x, cond)
    assert np.abs(agcm - g["agcm_out"]).max() <= 1e-4
    d = np.abs(out - g["out"])
    assert d.mean() <= 8e-3 and d.max() <= 8e-2


def test_int8_mixed_layer_accumulators_fixture_is_self_consistent():
    """The integer accumulators recorded next to the reference modules' outputs reproduce those outputs through the
    de-quantisation identity the kind::i8 epilogue uses: conv(x^, w) + b = acc * (s * ws) + b + z * ws * sum_valid(w)."""
    g = load_golden("int8mixed_layers_64x96.npz")
    raw = load_golden("weights_int8_mixed_qat.npz")
    for layer in [str(x) for x in g["layers"]]:
        q, acc, out = g[layer + "|q"][0].astype(np.float64), g[layer + "|acc"][0].astype(np.float64), g[layer + "|out"][0]
        w8 = raw[layer + ".weight_int8"].astype(np.float64)
        ws = raw[layer + ".w_scale"].astype(np.float64).reshape(-1, 1, 1)
        s, z = float(raw[layer + ".x_scale"]), float(raw[layer + ".x_zero"])
        stride = int(g[layer + "|stride"])
        ones = np.ones_like(q)
        wsum = O.conv2d(ones.astype(np.float32), w8.astype(np.float32), None, stride=stride).astype(np.float64)   # sum of w over VALID taps
        deq = acc * (s * ws) + raw[layer + ".bias"].astype(np.float64).reshape(-1, 1, 1) + z * ws * wsum
        assert np.abs(deq - out).max() <= 2e-4 * max(1.0, float(np.abs(out).max())), layer


# ---------------------------------------------------------------------------------------------------------------------
# HG stage: fixtures = the reference's own HG_Composite run on CPU with the seeded stand-in for the
# absent HG.pt (scripts/make_golden_hg.py)
# ---------------------------------------------------------------------------------------------------------------------
HG_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "hg_*.npz")))


@pytest.fixture(scope="module")
def hg_weights():
    from hdr_realtime_video_pipeline_b200.synth import hg_random_state_dict
    return hg_random_state_dict(0)


def test_hg_seeded_weights_have_the_reference_key_set(hg_weights):
    from hdr_realtime_video_pipeline_b200.synth import hg_state_dict_spec
    spec = hg_state_dict_spec()
    assert len(spec) == 92 and set(spec) == set(hg_weights)          # Hallucination_Generator().state_dict(): 92 entries
    assert spec["Up_conv1.0.weight"] == (2048, 512, 3, 3) and spec["conv6.weight"] == (512, 1024, 1, 1)
    again = __import__("hdr_realtime_video_pipeline_b200.synth", fromlist=["x"]).hg_random_state_dict(0)
    assert all(np.array_equal(again[k], hg_weights[k]) for k in spec)  # seed -> identical tensors


@pytest.mark.parametrize("name", HG_CASES)
def test_hg_stage_matches_reference(name, hg_weights):
    g = load_golden(name)
    assert np.array_equal(O.hg_mask(g["base_out"][0]).astype(np.uint8), g["mask"][0])     # threshold mask is exact
    out = O.hg_stage(hg_weights, g["base_out"])
    assert out.shape == g["hg_out"].shape
    assert np.abs(out - g["hg_out"]).max() <= 2e-6                   # fp32 summation order only
    folded = O.hg_stage(O.hg_fold_bn(hg_weights), g["base_out"])     # eval BatchNorm folded into the convs (FusedBN arithmetic)
    assert np.abs(folded - g["hg_out"]).max() <= 2e-6


def test_hg_reflect_pad_and_pool_primitives():
    x = np.arange(2 * 6 * 10, dtype=np.float32).reshape(2, 6, 10)
    p = O.max_pool2(x)
    assert p.shape == (2, 3, 5) and p[1, 2, 4] == x[1, 5, 9] and p[0, 0, 0] == x[0, 1, 1]
    # F.pad(mode="reflect") on the right / bottom: index W + i reads W - 2 - i (HG_Composite_arch.py:94-101)
    base = np.random.default_rng(0).random((1, 3, 40, 50)).astype(np.float32)
    img = np.pad(base[0], ((0, 0), (0, 24), (0, 14)), mode="reflect")
    assert img.shape == (3, 64, 64) and np.array_equal(img[:, 40 + 3, :50], base[0, :, 40 - 2 - 3, :])
    assert np.array_equal(img[:, :40, 50 + 5], base[0, :, :, 50 - 2 - 5])


# ---------------------------------------------------------------------------------------------------------------------
# Letterbox (gui_scaling.py:228-244): the oracle restates cv2.resize (third party, opencv-python 4.13 in this image);
# pinned against cv2 itself.  INTER_AREA: bit-exact with and without the Intel-IPP dispatch.  INTER_CUBIC: bit-exact
# against OpenCV's own code path (IPP off); the closed-source IPP routine the wheel uses by default differs by <= 1 code.
# ---------------------------------------------------------------------------------------------------------------------
LETTERBOX_CASES = [(216, 384, 192, 108), (270, 480, 192, 108), (200, 300, 192, 108), (108, 192, 384, 216), (100, 133, 217, 160),
                   (90, 160, 192, 108), (120, 160, 192, 108), (54, 96, 217, 123), (108, 192, 192, 108), (300, 200, 192, 108),
                   (324, 576, 192, 108)]


def _reference_letterbox(cv2, frame, out_w, out_h):
    """gui_scaling.py:228-244, restated around the real cv2.resize."""
    h, w = frame.shape[:2]
    if w == out_w and h == out_h:
        return frame
    scale = min(out_w / max(w, 1), out_h / max(h, 1))
    new_w = max(1, int(round(w * scale)))
    new_h = max(1, int(round(h * scale)))
    interp = cv2.INTER_AREA if scale < 1.0 else cv2.INTER_CUBIC
    resized = cv2.resize(frame, (new_w, new_h), interpolation=interp)
    canvas = np.zeros((out_h, out_w, 3), dtype=frame.dtype)
    x, y = (out_w - new_w) // 2, (out_h - new_h) // 2
    canvas[y:y + new_h, x:x + new_w] = resized
    return canvas


@pytest.mark.parametrize("case", LETTERBOX_CASES, ids=lambda c: "%dx%d_to_%dx%d" % c)
def test_letterbox_matches_cv2(case):
    cv2 = pytest.importorskip("cv2")
    h, w, out_w, out_h = case
    frame = np.random.default_rng(h * 7 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    mine = O.letterbox_bgr(frame, out_w, out_h)
    ipp = cv2.ipp.useIPP()
    try:
        cv2.ipp.setUseIPP(False)
        assert np.array_equal(mine, _reference_letterbox(cv2, frame, out_w, out_h))           # OpenCV's own arithmetic: bit-exact
        cv2.ipp.setUseIPP(True)
        ref = _reference_letterbox(cv2, frame, out_w, out_h)
        shrink = O.letterbox_geometry(h, w, out_h, out_w)[4]
        d = np.abs(mine.astype(np.int32) - ref.astype(np.int32))
        assert d.max() <= (0 if shrink else 1)                                              # IPP cubic: within one code
    finally:
        cv2.ipp.setUseIPP(ipp)


def test_letterbox_geometry_rounds_like_python():
    # int(round(x)) is round-half-even: 2.5 -> 2, 3.5 -> 4 (gui_scaling.py:235-236)
    assert O.letterbox_geometry(5, 10, 2, 100)[:2] == (2, 4)
    assert O.letterbox_geometry(1080, 1920, 1080, 1920)[:2] == (1080, 1920)
    assert O.letterbox_geometry(2160, 3840, 1080, 1920) == (1080, 1920, 0, 0, True)
    assert O.letterbox_geometry(1080, 1440, 1080, 1920) == (1080, 1440, 0, 240, False)
