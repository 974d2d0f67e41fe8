#!/usr/bin/env python
"""GPU bring-up diagnostics (run on the B200 box through gpurun).  Each stage runs in its own subprocess so a
trapped kernel (sticky CUDA error) does not take the later stages down with it.

    python scripts/gpu_diag.py            # all stages, log to gpurun_out/diag.log
    python scripts/gpu_diag.py selftest   # one stage in-process
"""
import json
import os
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

KINDS = {0: "NAT3x3", 1: "NAT1x1", 2: "NAT3x3_C8", 3: "NAT1x1_C8", 4: "PAR3x3S2", 5: "PAR1x1"}


def _mk(precision="fp16"):
    import hdr_realtime_video_pipeline_b200 as hb
    return hb.HDRTVNetB200(os.path.join(REPO, "tests/golden/weights_hr.npz"), precision=precision, warmup_passes=0,
                           use_hg=False, debug_library=True)


def stage_selftest(args):
    net = _mk()
    cases = [
        # kind, cin, cout, H, W, flags
        (1, 16, 16, 8, 128, 0), (1, 16, 32, 8, 128, 0), (1, 32, 64, 8, 128, 0), (1, 64, 64, 12, 200, 4),
        (1, 64, 16, 12, 200, 0), (1, 16, 128, 9, 130, 4),
        (3, 8, 64, 8, 128, 4),
        (0, 32, 32, 8, 128, 0), (0, 32, 32, 20, 300, 4), (0, 32, 32, 20, 300, 4 | 8), (0, 32, 32, 20, 300, 4 | 8 | 16),
        (0, 32, 128, 10, 140, 4 | 1), (0, 32, 128, 10, 140, 4 | 1 | 8 | 16), (0, 32, 3, 10, 140, 2),
        (2, 8, 64, 20, 300, 4), (2, 8, 32, 20, 300, 4 | 16),
        (4, 32, 32, 20, 300, 4), (4, 32, 32, 21, 301, 4 | 16), (4, 64, 64, 20, 300, 4), (4, 64, 16, 21, 301, 0),
        (5, 64, 64, 12, 200, 4), (5, 64, 64, 13, 201, 4),
        (0, 32, 32, 300, 700, 4),
    ]
    if args:
        cases = [c for c in cases if str(c[0]) in args]
    for kind, cin, cout, h, w, fl in cases:
        t0 = time.time()
        try:
            mx, ref = net.conv_selftest(kind, cin, cout, h, w, fl)
            ok = "OK " if mx < 2e-2 * max(ref, 1.0) else "BAD"
            print(f"selftest {ok} kind={KINDS[kind]:10s} cin={cin:3d} cout={cout:3d} {h}x{w} flags={fl:2d} "
                  f"max_abs={mx:.4g} ref_max={ref:.3g}  ({time.time() - t0:.2f}s)", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"selftest ERR kind={KINDS[kind]} cin={cin} cout={cout} {h}x{w} flags={fl}: {e}", flush=True)
            break


def _compare(name, a, b):
    import numpy as np
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    print(f"  {name:10s} max_abs={d.max():.3e} mean_abs={d.mean():.3e} ref_max={np.abs(b).max():.3g}", flush=True)
    return float(d.max())


def stage_net(args):
    import numpy as np
    import torch
    from oracle import hdrtvnet_oracle as O
    from hdr_realtime_video_pipeline_b200.synth import synth_frame
    precision = args[0] if args else "fp16"
    h, w = (int(args[1]), int(args[2])) if len(args) >= 3 else (72, 100)
    net = _mk(precision)
    sd = dict(np.load(os.path.join(REPO, "tests/golden/weights_hr.npz")))
    frame = synth_frame(0, h, w, "noise")
    npdt = np.float16 if precision == "fp16" else np.float32
    x_o, c_o = O.preprocess(frame, npdt)
    x, cond = net.preprocess(frame)
    torch.cuda.synchronize()
    print(f"net {precision} {h}x{w}")
    _compare("x", x.float().cpu().numpy(), x_o.astype(np.float32))
    _compare("cond", cond.float().cpu().numpy(), c_o.astype(np.float32))
    out, agcm = net.infer((x, cond))
    torch.cuda.synchronize()
    sdf = {k: np.asarray(v, np.float32) for k, v in sd.items()}
    xf, cf = x_o.astype(np.float32), c_o.astype(np.float32)
    fea = O.classifier(sdf, cf[0])
    agcm_o = O.agcm(sdf, xf[0], cf[0])
    out_o, inter = O.le(sdf, agcm_o, agcm_o, return_intermediates=True)
    dbg = net.debug_tensors()
    if "fea" in dbg:
        _compare("fea", dbg["fea"].reshape(-1), fea)
    _compare("agcm_out", agcm.float().cpu().numpy()[0], agcm_o)
    for k in ("cond", "cond1", "cond2", "cond3", "cond4", "fea0", "fea1", "fea2", "fea3"):
        if k in dbg:
            _compare(k, dbg[k], inter[k])
    _compare("out", out.float().cpu().numpy()[0], out_o)
    print("launches:", net.launch_count(), "workspace MB:", net.workspace_bytes() / 1e6, flush=True)


def stage_pack(args):
    import numpy as np
    import torch
    from oracle import hdrtvnet_oracle as O
    import hdr_realtime_video_pipeline_b200 as hb
    g = dict(np.load(os.path.join(REPO, "tests/golden/pack.npz")))
    net = _mk("fp16")
    for tag in ("32", "16"):
        t = torch.from_numpy(g["in" + tag]).cuda()
        fr = hb.tensor_to_rgb48_bytes(t, {})
        got = fr.numpy().copy()
        fr.release()
        print("rgb48", tag, "bit-exact:", np.array_equal(got, g["rgb48_" + tag]), flush=True)
        bgr = net.postprocess(t.clone()).copy()
        print("bgr24", tag, "bit-exact:", np.array_equal(bgr, g["bgr24_" + tag]), flush=True)


STAGES = {"selftest": stage_selftest, "net": stage_net, "pack": stage_pack}

if __name__ == "__main__":
    if len(sys.argv) > 1:
        STAGES[sys.argv[1]](sys.argv[2:])
        sys.exit(0)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(REPO, "gpurun_out", "diag.log"), "w")
    plan = [["pack"], ["selftest", "1"], ["selftest", "3"], ["selftest", "0"], ["selftest", "2"], ["selftest", "4"],
            ["selftest", "5"], ["net", "fp32", "72", "100"], ["net", "fp16", "72", "100"], ["net", "fp16", "136", "248"]]
    for st in plan:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), *st], capture_output=True, text=True, timeout=240)
            txt = r.stdout + ("\n[stderr]\n" + r.stderr[-3000:] if r.returncode != 0 else "")
            rc = r.returncode
        except subprocess.TimeoutExpired as e:
            txt, rc = f"TIMEOUT {e}", -9
        msg = f"===== {' '.join(st)} rc={rc} ({time.time() - t0:.1f}s)\n{txt}\n"
        print(msg, flush=True)
        log.write(msg)
        log.flush()
