#!/usr/bin/env python
"""bench.py — HDRTVNet++ per-frame SDR->HDR inference throughput on B200 (driver contract in the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload both|1080p|4k|540p] [--precision fp16|fp32|int8-mixed|int8-full]
                    [--impl b200|reference]

A "step" is one frame through the hot path: uint8 BGR24 -> normalise -> AGCM -> LE -> RGB48 pack.

Default run (N = 1): BOTH BASELINE workloads in one JSON line —
  workloads["4k"]    BASELINE configs[2]: 3840x2160 FP16, fused RGB48 pack into the pinned host ring   (the north-star target,
                     >= 60 frames/s; the top-level value / e2e / roofline of the line are this workload's, at every N)
  workloads["1080p"] BASELINE configs[1]: 1920x1080 FP16 batch-1 playback stream, incl. p50 / p99 frame latency
N > 1 (torchrun, one process per GPU): BASELINE configs[3], the frame-sharded 4K export — every rank runs
  hdr_realtime_video_pipeline_b200.export_clip over its contiguous chunk of an (N x K)-frame clip (weak scaling: K frames per
  GPU; no collective on the pixel path, one all_gather of the per-rank descriptor records at the end); the merged
  descriptors are checked against the checksums of a serial single-GPU pass.

  value : frames/s with the input frames already resident in HBM (device u8 in, device u16 out), CUDA events, max over ranks
  e2e   : frames/s through the public API with HOST buffers (pinned u8 frame in, RGB48 in a pinned ring slot out, H2D and D2H
          inside the timed region), wall clock between device synchronisations, max over ranks
  gpu_eager_baseline : the UNMODIFIED reference (baseline/_ref, scripts/install_reference.py) on the same GPU: HDRTVNetTorch(
          device="cuda", precision="fp16") eager / cuDNN channels_last + the reference feeder pack — the bar to beat
  cpu_baseline / --impl reference : the reference's own CPU eager path (setup_cpu), all host threads, bounded sample
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {"540p": (540, 960), "1080p": (1080, 1920), "4k": (2160, 3840)}
CONFIG_NAME = {
    "540p": "HDRTVNet++ (AGCM+LE) 960x540 synthetic BT.709 SDR frames (BASELINE configs[0] size)",
    "1080p": "HDRTVNet++ FP16 1920x1080 single-B200 real-time playback stream (batch 1, latency-bound) — BASELINE configs[1]",
    "4k": "HDRTVNet++ FP16 3840x2160 single-B200 inference with fused RGB48 pack into pinned host ring — BASELINE configs[2]",
    "4k_export": "HDRTVNet++ FP16 4K export of a synthetic clip frame-sharded across the GPUs (K frames per GPU, 2400 at "
                 "--steps 300 x 8) — BASELINE configs[3]",
}
FLOP_PER_PX = 221054.0          # SURVEY §8d / BASELINE.md §2: 2 x 110 527 conv+linear MACs per pixel
# algorithmic MACs per full-resolution pixel of the largest launches
KERNEL_MAC_PER_PX = {
    "LE.cond_chain+sft0.L0": 9920 + 9216 + 1024,         # cond_first + CondNet1 + stage 0 of SFT_layer1/2 (16 -> 64)
    "LE.CondNet{2,3,4}.0": 27648,                         # three 3x3 s2 64->64 convs on cond
    "AGCM.chain": 4480,
    "LE.up_conv3.0": 9216 + 1536,                         # 3x3 32->128 at H/2 + SFT_layer2 stage 1
    "LE.conv_first+HR_conv1": 864 + 1536 + 9216,          # + SFT_layer1 stage 1
    "LE.HR_conv2+conv_last": 9216 + 864,
    "LE.recon_trunk1.0.conv1+conv2": 2304 + 2304 + 384, "LE.recon_trunk5.0.conv1+conv2": 2304 + 2304 + 384,
    "LE.HR_conv1": 9216, "LE.HR_conv2": 9216, "LE.conv_first": 864 + 1536, "LE.conv_last": 864,
    "LE.down_conv1": 2304 + 384,
}
# whole-frame DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum of every launch of one frame) from the committed ncu
# launch lists: profiles/r2_traffic.json (this round) else profiles/r1_ncu_top_kernels.json
NCU_FRAME_TRAFFIC, NCU_KERNEL_TRAFFIC, NCU_TRAFFIC_SOURCE = {}, {}, None
for _name in ("r2_traffic.json", "r1_ncu_top_kernels.json"):
    try:
        with open(os.path.join(REPO, "profiles", _name)) as _f:
            _j = json.load(_f)
        for _wl in ("1080p", "4k"):
            if _j.get(f"frame_dram_bytes_{_wl}", {}).get("total"):
                NCU_FRAME_TRAFFIC[_wl] = int(_j[f"frame_dram_bytes_{_wl}"]["total"])
            NCU_KERNEL_TRAFFIC[_wl] = {k: int(v) for k, v in _j.get(f"dram_bytes_per_launch_{_wl}", {}).items()}
        NCU_TRAFFIC_SOURCE = "profiles/" + _name
        break
    except Exception:
        continue
NCU_HG_TRAFFIC = {}        # DRAM bytes of one HG pass (ncu launch list of scripts/profile_hg.py, profiles/r2_hg_traffic.json)
try:
    with open(os.path.join(REPO, "profiles", "r2_hg_traffic.json")) as _f:
        _j = json.load(_f)
    for _wl in ("1080p", "4k"):
        if _j.get(f"frame_dram_bytes_{_wl}", {}).get("total"):
            NCU_HG_TRAFFIC[_wl] = int(_j[f"frame_dram_bytes_{_wl}"]["total"])
except Exception:
    pass
WEIGHTS = os.path.join(REPO, "tests", "golden", "weights_hr.npz")
# the reference's INT8 checkpoints (src/models/weights/original/pytorch_int8/hr/*_qat.pt) as raw-array fixtures
W_INT8 = {"int8-full": os.path.join(REPO, "tests", "golden", "weights_int8_full_qat.npz"),
          "int8-mixed": os.path.join(REPO, "tests", "golden", "weights_int8_mixed_qat.npz")}


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops": float(p.get("bf16_tflops_sustained", 1381.4)), "tflops_burst": float(p.get("bf16_tflops", 1659.2)),
                "hbm_gbs": float(p.get("hbm_gbs", 6556.2)),
                "source": "MEASURED_PEAKS.json (cuBLAS bf16 sustained / burst, STREAM-style copy)"}
    return {"tflops": 1400.0, "tflops_burst": 1650.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, sampled through NVML every ~5 ms (the timed region of a
    20-step 1080p run is 20 ms: nvidia-smi's 100 ms loop cannot see it)."""
    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
               ("hw_power_brake_slowdown", 0x80), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index: int, period_s: float = 0.005):
        self.idx, self.period, self.rows, self._stop, self._th, self.h = gpu_index, period_s, [], threading.Event(), None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].strip().isdigit() else gpu_index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _run(self):
        nv, h = self.nv, self.h
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                why = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                mw = nv.nvmlDeviceGetPowerUsage(h)
                self.rows.append((time.perf_counter(), float(mhz), int(why), mw / 1000.0))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.h is None:
            return
        self._stop.clear()
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()

    def stop(self, t0, t1):
        if self.h is None:
            return None
        self._stop.set()
        self._th.join(timeout=1.0)
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or [r for r in self.rows if t0 - 0.05 <= r[0] <= t1 + 0.05]
        self.rows = []
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        bits = 0
        for r in rows:
            bits |= r[2]
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for n, b in self.REASONS if bits & b), "samples": len(rows),
                "power_w_max": float(max(r[3] for r in rows)), "sampling": "NVML, 5 ms period, samples inside the timed region"}


# ---------------------------------------------------------------------------------------------- reference legs
def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs are meant to use every core this process may run on."""
    import torch
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


class CpuReference:
    """The reference's setup_cpu configuration: the unmodified reference when baseline/_ref is installed
    (HDRTVNetTorch(device="cpu", precision="fp32", compile_model=False) -> preprocess / infer / feeder pack), else its
    torch-CPU restatement oracle/torch_port.py."""

    def __init__(self):
        from oracle import reference_loader as RL
        self.ref = RL.load()
        if self.ref is not None:
            with contextlib.redirect_stdout(sys.stderr):
                self.net = self.ref.HDRTVNetTorch(self.ref.weights("HR.pt"), device="cpu", precision="fp32", compile_model=False,
                                                  use_hg=False, warmup_passes=0)
            self.kind = "reference"
            self.what = ("unmodified reference (baseline/_ref): HDRTVNetTorch(device='cpu', precision='fp32', compile_model=False)"
                         ".preprocess/.infer + _tensor_to_rgb48_bytes")
            self.state = {}
        else:
            from oracle import torch_port as TP
            self.TP = TP
            self.sd = TP.to_torch_state(dict(np.load(WEIGHTS)))
            self.kind = "port"
            self.what = "oracle/torch_port.py (torch CPU eager fp32 restatement of the reference's setup_cpu path)"

    def frame(self, f):
        if self.ref is None:
            return self.TP.process_rgb48(self.sd, f)
        import torch
        with torch.inference_mode():
            out = self.net.infer(self.net.preprocess(f))
            return self.ref.feeders._tensor_to_rgb48_bytes(out, self.state)


def cpu_reference_sample(height: int, width: int, budget_s: float = 22.0, max_frames: int = 12):
    import torch
    from hdr_realtime_video_pipeline_b200.synth import synth_frame
    _use_all_host_threads()
    cpu = CpuReference()
    cpu.frame(synth_frame(0, height, width))                      # first frame discarded
    times, t_start = [], time.perf_counter()
    for i in range(max_frames):
        f = synth_frame(i + 1, height, width)
        t0 = time.perf_counter()
        cpu.frame(f)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s and len(times) >= 2:
            break
    return times, torch.get_num_threads(), cpu


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    _use_all_host_threads()
    wl = "4k" if args.workload == "both" else args.workload
    h, w = WORKLOADS[wl]
    sh, sw = 540, 960                                           # bounded sample: one config-1-size frame per step
    from hdr_realtime_video_pipeline_b200.synth import synth_frame
    cpu = CpuReference()
    for i in range(max(1, args.warmup)):
        cpu.frame(synth_frame(i, sh, sw))
    t0 = time.perf_counter()
    for i in range(args.steps):
        cpu.frame(synth_frame(100 + i, sh, sw))
    dt = time.perf_counter() - t0
    px_per_s = args.steps * sh * sw / dt
    fps = px_per_s / (h * w)
    cores = torch.get_num_threads()
    sample = (f"each step = one {sw}x{sh} frame (1/{(h * w) // (sh * sw)} of a {w}x{h} frame) through {cpu.what}; frames/s scaled "
              f"by pixel count to {w}x{h}")
    line = {"impl": "reference", "metric": "HDRTVNet++ frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1000.0, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIG_NAME[wl], "height": h, "width": w},
            "pixels_per_s": px_per_s,
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "host_cpu_count": os.cpu_count(), "kind": cpu.kind,
                             "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=JSON_OUT, flush=True)


def gpu_eager_baseline(dev, h, w, n_frames, host_frames):
    """The unmodified reference on the same GPU (SURVEY §8d "same-box GPU reference"): HDRTVNetTorch(device='cuda',
    precision='fp16') = eager cuDNN, channels_last, cudnn.benchmark, no torch.compile on NVIDIA (hdrtvnet_torch.py:
    1573-1607) -> preprocess -> infer -> the reference feeder's _tensor_to_rgb48_bytes, host frame in, pinned RGB48 out."""
    import torch
    from oracle import reference_loader as RL
    ref = RL.load()
    if ref is None:
        return {"unavailable": "baseline/_ref not installed (scripts/install_reference.py)"}
    if dev.index not in (None, 0):
        return {"unavailable": "the reference wrapper always runs on cuda:0 (hdrtvnet_torch.py:1682)"}
    try:
        with contextlib.redirect_stdout(sys.stderr):
            net = ref.HDRTVNetTorch(ref.weights("HR.pt"), device="cuda", precision="fp16", compile_model=False, use_hg=False,
                                    warmup_passes=0)
        state = {}

        def one(i):
            with torch.inference_mode():
                out = net.infer(net.preprocess(host_frames[i % len(host_frames)]))
            payload = ref.feeders._tensor_to_rgb48_bytes(out, state)
            if hasattr(payload, "release"):
                payload.wait_ready()
                payload.release()

        for i in range(4):
            one(i)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for i in range(n_frames):
            one(i)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        # infer-only device time (CUDA events), the part this repo's kernels replace
        x, c = net.preprocess(host_frames[0])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(max(3, n_frames // 3)):
            with torch.inference_mode():
                net.infer((x, c))
        e1.record()
        torch.cuda.synchronize(dev)
        infer_ms = e0.elapsed_time(e1) / max(3, n_frames // 3)
        del net
        torch.cuda.empty_cache()
        return {"value": n_frames / dt, "unit": "frames/s", "ms_per_frame": dt / n_frames * 1000.0, "infer_only_ms": infer_ms,
                "frames": n_frames, "kind": "reference",
                "what": "unmodified reference (baseline/_ref) on this GPU: HDRTVNetTorch(device='cuda', precision='fp16', "
                        "compile_model=False) preprocess -> infer -> _tensor_to_rgb48_bytes, host u8 frame in, pinned RGB48 out, "
                        f"torch {torch.__version__} eager cuDNN channels_last"}
    except Exception as exc:                                    # a baseline must never take the product line down
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}


# ---------------------------------------------------------------------------------------------- B200 arm
def measure_workload(net, packer, hb, torch, dev, barrier, wl, K, Wm, rank, local_rank, first_frame, precision, sampler):
    """Device-resident throughput, end-to-end throughput (one-call and three-call API), batch-1 latency, K1 / K8 bandwidth and
    the per-launch times of one workload on this rank."""
    h, w = WORKLOADS[wl]
    px = h * w
    n_distinct = 32 if px <= 1080 * 1920 else 16                 # >= 199 MB of distinct input frames (> 126 MB L2)
    host_frames = [torch.from_numpy(hb.synth_frame(first_frame + i, h, w)).pin_memory() for i in range(n_distinct)]
    dev_frames = [f.to(dev) for f in host_frames]
    out_dev = torch.empty((h, w, 3), dtype=torch.uint16, device=dev)
    res = {}

    def step_device(i):
        x, c = net.preprocess_device(dev_frames[i % n_distinct], assume_ready=True)
        out = net.infer((x, c))
        packer.pack_device(out, out_dev)
        return out

    for i in range(Wm):
        step_device(i)
    barrier()
    if rank == 0:
        sampler.start()
        time.sleep(0.03)
    l0 = net.launch_count() + packer.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    infer_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record()
    for i in range(K):
        x, c = net.preprocess_device(dev_frames[i % n_distinct], assume_ready=True)
        infer_ev[i][0].record()
        out = net.infer((x, c))
        infer_ev[i][1].record()
        packer.pack_device(out, out_dev)
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    res["dev_ms"] = ev0.elapsed_time(ev1)
    res["launches"] = net.launch_count() + packer.launch_count() - l0
    res["clocks"] = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    res["infer_ms"] = float(np.mean([a.elapsed_time(b) for a, b in infer_ev]))

    # ---- end to end through the public API, host buffers --------------------------------------------
    state = {}
    host_np = [f.numpy() for f in host_frames]

    def e2e_run(submit, in_flight):
        pending, checks = [], []
        for i in range(Wm):
            submit(i).release()
        barrier()
        e0 = time.perf_counter()
        for i in range(K):
            pending.append(submit(i))
            if len(pending) >= in_flight:                            # consumer side of the ring: wait + release in order
                fr = pending.pop(0)
                checks.append(int(fr.numpy()[h // 2, w // 2, 0]))
                fr.release()
        for fr in pending:
            checks.append(int(fr.numpy()[h // 2, w // 2, 0]))
            fr.release()
        barrier()
        return time.perf_counter() - e0, checks

    e2e_s, checks = e2e_run(lambda i: net.process_rgb48(host_np[i % n_distinct]), 3)
    e2e3_s, checks3 = e2e_run(lambda i: hb.tensor_to_rgb48_bytes(net.infer(net.preprocess(host_np[i % n_distinct])), state), 2)
    if checks != checks3:
        bad = [(i, a, b) for i, (a, b) in enumerate(zip(checks, checks3)) if a != b]
        raise RuntimeError(f"one-call and three-call frame paths disagree on {len(bad)} of {len(checks)} frames: {bad[:8]}")
    res["e2e_s"], res["e2e3_s"], res["checks"] = e2e_s, e2e3_s, checks

    # ---- batch-1 latency (config 2 is latency-bound): serial frames, host-timed submit .. RGB48 slot ready --------
    def latency_run(submit):
        lat = []
        for i in range(5):                                           # warm-up: the first frames after a mode switch are not steady state
            submit(i).release()
        for i in range(max(30, min(K, 200))):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            fr = submit(i)
            fr.wait_ready()
            lat.append((time.perf_counter() - t0) * 1000.0)
            fr.release()
        return lat

    res["lat"] = latency_run(lambda i: net.process_rgb48(host_np[i % n_distinct], serial=True))
    res["lat3"] = latency_run(lambda i: hb.tensor_to_rgb48_bytes(net.infer(net.preprocess(host_np[i % n_distinct])), state))

    # ---- K1 (normalise + condition image) and K8 (RGB48 pack) alone: achieved GB/s over their algorithmic bytes ----
    if rank == 0:
        reps = 40
        esz = 2 if net._dtype == torch.float16 else 4
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pipeline = net._pipeline
        net._pipeline = False                                        # preprocess alone, on the current stream
        for i in range(3):
            net.preprocess_device(dev_frames[i % n_distinct], assume_ready=True)
        e0.record()
        for i in range(reps):
            net.preprocess_device(dev_frames[i % n_distinct], assume_ready=True)
        e1.record()
        torch.cuda.synchronize(dev)
        k1_ms = e0.elapsed_time(e1) / reps
        net._pipeline = pipeline
        out = step_device(0)
        e0.record()
        for i in range(reps):
            packer.pack_device(out, out_dev)
        e1.record()
        torch.cuda.synchronize(dev)
        k8_ms = e0.elapsed_time(e1) / reps
        k1_bytes = px * (3 + 3 * esz) + (h // 4) * (w // 4) * 3 * esz
        k8_bytes = px * (3 * esz + 6)
        res["k1"] = {"ms": k1_ms, "algorithmic_bytes": k1_bytes, "achieved_GBps": k1_bytes / (k1_ms * 1e-3) / 1e9,
                     "kernels": "normalize_vec16_kernel + cond_aa_kernel (hdrtv_preprocess): 3 B/px in, planar x + 1/16-size cond out"}
        res["k8"] = {"ms": k8_ms, "algorithmic_bytes": k8_bytes, "achieved_GBps": k8_bytes / (k8_ms * 1e-3) / 1e9,
                     "kernels": "pack_rgb48_kernel (hdrtv_pack_rgb48): planar out in, 6 B/px RGB48 out",
                     "note": "the same frame is packed repeatedly here: its 12-37 MB fit the 126 MB L2, so this is an L2-resident figure"}

    # ---- per-launch device times of one frame (CUDA events between launches, median of 5) -> top-kernel rooflines
    res["top_kernels"] = []
    if rank == 0 and net._dtype == torch.float16:
        x, c = net.preprocess_device(dev_frames[0], assume_ready=True)
        runs = [net.time_plan((x, c)) for _ in range(5)]
        names = [n for n, _ in runs[0]]
        med = np.median(np.array([[t for _, t in r] for r in runs]), axis=0)
        for i in np.argsort(-med)[:5]:
            name = names[i].split(" ")[0]
            mac = KERNEL_MAC_PER_PX.get(name)
            entry = {"launch": names[i], "ms": float(med[i]), "share_of_infer": float(med[i] / med.sum())}
            if mac is not None:
                entry.update({"algorithmic_mac_per_px": mac, "achieved_tflops": 2.0 * mac * px / (med[i] / 1000.0) / 1e12})
            traffic = NCU_KERNEL_TRAFFIC.get(wl, {}).get(name)
            if traffic is not None:      # DRAM bytes of this launch (ncu, profiles/) over its CUDA-event time here
                entry["traffic_bytes_per_launch"] = traffic
                entry["hbm_GBps"] = traffic / (med[i] * 1e-3) / 1e9
            res["top_kernels"].append(entry)
    res["host_np"], res["n_distinct"] = host_np, n_distinct
    return res


# HG stage: conv shapes of Hallucination_Generator (nf = 64) -> FLOPs of one frame at the padded size
HG_CONVS = [  # (name, cin, cout, ksize, level)
    ("conv1.0", 3, 64, 3, 0), ("conv2.0", 64, 128, 3, 1), ("conv3_1.0", 128, 256, 3, 1), ("conv3_2.0", 256, 256, 3, 2),
    ("conv4_1.0", 256, 512, 3, 2), ("conv4_2.0", 512, 512, 3, 3), ("conv5_1.0", 512, 512, 3, 3), ("conv5_2.0", 512, 512, 3, 4),
    ("conv_code1.0", 512, 512, 3, 4), ("conv_code2.0", 512, 512, 3, 5), ("Up_conv1.0", 512, 2048, 3, 5), ("conv6", 1024, 512, 1, 4),
    ("Up_conv2.0", 512, 2048, 3, 4), ("conv7", 1024, 256, 1, 3), ("Up_conv3.0", 256, 1024, 3, 3), ("conv8", 512, 128, 1, 2),
    ("Up_conv4.0", 128, 512, 3, 2), ("conv9", 256, 64, 1, 1), ("Up_conv5.0", 64, 256, 3, 1), ("conv10", 128, 3, 1, 0),
    ("conv_last", 6, 3, 1, 0)]


def hg_flops(h, w):
    hp, wp = (h + 31) // 32 * 32, (w + 31) // 32 * 32
    per = {n: 2.0 * ci * co * k * k * (hp >> lv) * (wp >> lv) for n, ci, co, k, lv in HG_CONVS}
    return per, float(sum(per.values()))


def measure_hg(hb, torch, dev, wl, K, Wm, peaks, weights, with_reference, with_cpu=False):
    """AGCM + LE + HG (HG_Composite, the full HDRTVNet++): device-resident and end-to-end frames/s, the HG stage's own
    roofline (tensor-bound: 64..2048-channel convs on the K-streamed tcgen05 kernel) with per-launch TFLOP/s, and the
    unmodified reference with the same seeded HG weights on the same GPU (CUDA FP16 eager)."""
    from hdr_realtime_video_pipeline_b200.synth import hg_random_state_dict
    h, w = WORKLOADS[wl]
    px = h * w
    hg_sd = hg_random_state_dict(0)
    net = hb.HDRTVNetB200(weights, device=f"cuda:{dev.index}", precision="fp16", warmup_passes=0, use_hg=True, hg_weights=hg_sd)
    packer = hb.RGB48Packer(dev, ring_frames=3)
    n_distinct = 16 if px <= 1080 * 1920 else 8
    host_frames = [torch.from_numpy(hb.synth_frame(i, h, w)).pin_memory() for i in range(n_distinct)]
    dev_frames = [f.to(dev) for f in host_frames]
    host_np = [f.numpy() for f in host_frames]
    out_dev = torch.empty((h, w, 3), dtype=torch.uint16, device=dev)

    def step(i):
        out = net.infer(net.preprocess_device(dev_frames[i % n_distinct], assume_ready=True))
        packer.pack_device(out, out_dev)

    # The headline numbers of this workload are DENSE: the U-Net runs on every frame (HDRTV_HG_EARLY_OUT=0).  The product's
    # default skips it on frames without a pixel inside the highlight mask (three of the four synthetic content classes);
    # that figure is reported separately below as `mask_early_out`.
    os.environ["HDRTV_HG_EARLY_OUT"] = "0"
    for i in range(Wm):
        step(i)
    torch.cuda.synchronize(dev)
    l0 = net.launch_count() + packer.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step(i)
    e1.record()
    torch.cuda.synchronize(dev)
    step_ms = e0.elapsed_time(e1) / K
    launches = net.launch_count() + packer.launch_count() - l0
    os.environ["HDRTV_HG_EARLY_OUT"] = "1"
    for i in range(Wm):
        step(i)
    torch.cuda.synchronize(dev)
    e0.record()
    for i in range(K):
        step(i)
    e1.record()
    torch.cuda.synchronize(dev)
    gated_ms = e0.elapsed_time(e1) / K
    os.environ["HDRTV_HG_EARLY_OUT"] = "0"
    # end to end: host u8 frame in -> RGB48 in a pinned ring slot, one C-ABI call per frame, three frames in flight
    pending = []
    for i in range(Wm):
        net.process_rgb48(host_np[i % n_distinct]).release()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for i in range(K):
        pending.append(net.process_rgb48(host_np[i % n_distinct]))
        if len(pending) >= 3:
            fr = pending.pop(0)
            fr.wait_ready()
            fr.release()
    for fr in pending:
        fr.wait_ready()
        fr.release()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    # the stage alone + per-launch times
    out, _ = net.infer(net.preprocess_device(dev_frames[3 % n_distinct], assume_ready=True))
    base = net._gpu_out.clone()
    per, total = hg_flops(h, w)
    runs = [net.hg_time_plan(base) for _ in range(3)]
    names = [n for n, _ in runs[0]]
    med = np.median(np.array([[t for _, t in r] for r in runs]), axis=0)
    top = []
    for i in np.argsort(-med)[:5]:
        f = per[names[i].split()[0][3:]]
        tf = f / (med[i] * 1e-3) / 1e12
        top.append({"launch": names[i], "ms": float(med[i]), "achieved_tflops": tf, "tensor_frac_of_sustained_peak": tf / peaks["tflops"],
                    "tensor_frac_of_burst_peak": tf / peaks["tflops_burst"]})
    for _ in range(3):
        net.hg_stage(base)
    e0.record()
    for _ in range(10):
        net.hg_stage(base)
    e1.record()
    torch.cuda.synchronize(dev)
    stage_ms = e0.elapsed_time(e1) / 10
    stage_tf = total / (stage_ms * 1e-3) / 1e12
    res = {
        "config": f"HDRTVNet++ with the HG stage (HG_Composite: AGCM + LE + Hallucination_Generator) FP16 {w}x{h}, seeded stand-in "
                  "for the absent HG.pt (synth.hg_random_state_dict), HR.pt base",
        "value": 1000.0 / step_ms, "unit": "frames/s", "ms_per_step": step_ms, "steps": K, "gpu_launches": int(launches),
        "e2e": {"value": K / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": px * 3, "d2h_bytes_per_step": px * 6,
                "api": "HDRTVNetB200.process_rgb48 with HG weights installed (hdrtv_process_ex: LE -> hdrtv_hg -> fp32 RGB48 pack)"},
        "mask_early_out": {"value": 1000.0 / gated_ms, "unit": "frames/s", "ms_per_step": gated_ms,
                           "what": "the product default (HDRTV_HG_EARLY_OUT=1): frames without a pixel inside the highlight mask "
                                   "(max_c(base) <= 0.775) skip the U-Net through a device-side gate, bit-identical output; of the "
                                   "synthetic content classes only white+salt (every fourth frame) has highlights. Not the headline."},
        "roofline": {"bound": "tensor", "kernel": "gconv_kernel family of the HG stage (19 launches) + stage-in + tail, CUDA events "
                                                  "around 10 passes of hdrtv_hg on the launching stream",
                     "achieved": stage_tf, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": stage_tf / peaks["tflops"],
                     "frac_of_burst_peak": stage_tf / peaks["tflops_burst"], "ms_per_frame": stage_ms,
                     "algorithmic": f"{total / px:.0f} FLOP per output pixel x {px} px ({total / 1e12:.3f} TFLOP per frame at the "
                                    "padded size)", "peak_source": peaks["source"],
                     "traffic": NCU_HG_TRAFFIC.get(wl), "top_kernels": top},
    }
    os.environ.pop("HDRTV_HG_EARLY_OUT", None)
    packer.close()
    net.close()
    del net, packer
    torch.cuda.empty_cache()
    if with_reference:
        res["gpu_eager_baseline"] = gpu_eager_baseline_hg(dev, host_np, hg_sd, torch)
    if with_cpu:
        res["cpu_baseline"] = cpu_baseline_hg(hg_sd, h, w, torch)
    return res


def cpu_baseline_hg(hg_sd, h, w, torch):
    """The reference's CPU eager path WITH the HG stage (setup_cpu + HG_Composite) on a bounded sample: 960x540 frames, all
    host threads, first frame discarded; frames/s scaled by pixel count to the workload."""
    from oracle import reference_loader as RL
    from hdr_realtime_video_pipeline_b200.synth import synth_frame
    ref = RL.load()
    if ref is None:
        return {"unavailable": "baseline/_ref not installed (scripts/install_reference.py)"}
    try:
        import tempfile
        _use_all_host_threads()
        path = os.path.join(tempfile.mkdtemp(), "HG.pt")
        torch.save({k: torch.from_numpy(np.asarray(v)) for k, v in hg_sd.items()}, path)
        with contextlib.redirect_stdout(sys.stderr):
            net = ref.HDRTVNetTorch(ref.weights("HR.pt"), device="cpu", precision="fp32", compile_model=False, use_hg=True,
                                    hg_weights=path, warmup_passes=0)
        sh, sw = 540, 960
        times = []
        t_start = time.perf_counter()
        for i in range(6):
            f = synth_frame(i, sh, sw)
            t0 = time.perf_counter()
            with torch.inference_mode():
                net.postprocess(net.infer(net.preprocess(f)))
            if i:
                times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_start > 20.0 and len(times) >= 2:
                break
        px_s = sh * sw / float(np.mean(times))
        return {"value": px_s / (h * w), "unit": "frames/s", "cores": torch.get_num_threads(), "host_cpu_count": os.cpu_count(),
                "kind": "reference", "pixels_per_s": px_s,
                "sample": f"{len(times)} synthetic {sw}x{sh} frames through the unmodified reference (HDRTVNetTorch device='cpu', "
                          f"fp32, use_hg=True -> HG_Composite), first frame discarded, {float(np.mean(times)) * 1000:.0f} ms/frame; "
                          f"frames/s scaled by pixel count to {w}x{h}"}
    except Exception as exc:
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}


def gpu_eager_baseline_hg(dev, host_np, hg_sd, torch):
    """The unmodified reference with HG (HG_Composite through HDRTVNetTorch, CUDA FP16 eager, cudnn.benchmark) on this GPU."""
    from oracle import reference_loader as RL
    ref = RL.load()
    if ref is None:
        return {"unavailable": "baseline/_ref not installed (scripts/install_reference.py)"}
    try:
        import tempfile
        path = os.path.join(tempfile.mkdtemp(), "HG.pt")
        torch.save({k: torch.from_numpy(np.asarray(v)) for k, v in hg_sd.items()}, path)
        with contextlib.redirect_stdout(sys.stderr):
            net = ref.HDRTVNetTorch(ref.weights("HR.pt"), device="cuda", precision="fp16", compile_model=False, use_hg=True,
                                    hg_weights=path, warmup_passes=0)
        state = {}

        def one(i):
            with torch.inference_mode():
                out = net.infer(net.preprocess(host_np[i % len(host_np)]))
            payload = ref.feeders._tensor_to_rgb48_bytes(out, state)
            if hasattr(payload, "release"):
                payload.wait_ready()
                payload.release()

        for i in range(3):
            one(i)
        torch.cuda.synchronize(dev)
        n = 12 if host_np[0].shape[0] > 1080 else 24
        t0 = time.perf_counter()
        for i in range(n):
            one(i)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        del net
        torch.cuda.empty_cache()
        return {"value": n / dt, "unit": "frames/s", "ms_per_frame": dt / n * 1000.0, "frames": n, "kind": "reference",
                "what": "unmodified reference (baseline/_ref): HDRTVNetTorch(device='cuda', precision='fp16', use_hg=True) -> "
                        "HG_Composite eager cuDNN, preprocess -> infer -> _tensor_to_rgb48_bytes, same seeded HG weights"}
    except Exception as exc:
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}


def workload_entry(wl, r, world, K, peaks, precision):
    h, w = WORKLOADS[wl]
    px = h * w
    fps = world * K / (r["dev_ms"] / 1000.0)
    step_ms = r["dev_ms"] / K
    achieved_tf = FLOP_PER_PX * px / (step_ms / 1000.0) / 1e12
    traffic = NCU_FRAME_TRAFFIC.get(wl) if precision == "fp16" else None
    lat = np.array(r["lat"])
    fps_samples = np.sort(1000.0 / lat)
    low1 = float(np.mean(fps_samples[:max(1, len(fps_samples) // 100)]))
    e = {
        "config": CONFIG_NAME[wl], "height": h, "width": w,
        "value": fps, "unit": "frames/s", "ms_per_step": step_ms, "pixels_per_s": fps * px,
        "e2e": {"value": world * K / r["e2e_s"], "unit": "frames/s", "h2d_bytes_per_step": px * 3, "d2h_bytes_per_step": px * 6,
                "api": "HDRTVNetB200.process_rgb48 (one C-ABI call per frame: hdrtv_process_ex), pinned u8 frame in, RGB48 in a "
                       "pinned ring slot out, three frames in flight",
                "three_call_api": {"value": world * K / r["e2e3_s"], "unit": "frames/s",
                                   "api": "preprocess -> infer -> tensor_to_rgb48_bytes (the reference's call sequence)"},
                "timing": "wall clock between device synchronisations, includes host-side launch cost"},
        "latency_ms": {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)),
                       "one_percent_low_fps": low1, "frames": int(lat.size),
                       "three_call_api_p50": float(np.percentile(r["lat3"], 50)),
                       "what": "pinned u8 in -> RGB48 in pinned ring slot ready, one frame in flight (process_rgb48 serial=True); "
                               "'1 % low' = mean of the lowest 1 % of per-frame fps samples (main.py:599-604)"},
        "gpu_launches": int(r["launches"]),
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peaks["tflops"], "unit": "TFLOP/s",
                     "frac": achieved_tf / peaks["tflops"], "frac_of_burst_peak": achieved_tf / peaks["tflops_burst"],
                     "traffic": traffic,
                     "traffic_note": f"DRAM read + write bytes of all launches of one frame (ncu, {NCU_TRAFFIC_SOURCE}): inter-layer "
                                     "activations; the frame's algorithmic I/O is 9 B/px",
                     "hbm": ({"achieved_GBps": traffic / (step_ms * 1e-3) / 1e9, "peak_GBps": peaks["hbm_gbs"],
                              "frac": traffic / (step_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]} if traffic else None),
                     "kernel": "whole hot path of one frame (chain_p8_kernel, conv2x_p8_kernel, conv_p8_kernel family, classifier, "
                               "front end, pack); CUDA events around the timed steps on the launching stream",
                     "algorithmic": f"{FLOP_PER_PX:.0f} FLOP/px x {px} px per frame", "peak_source": peaks["source"],
                     "ms_per_frame": step_ms, "infer_only_ms": r["infer_ms"],
                     "note": "tcgen05.mma M=128 K=16 occupies the pipe for max(N/2, 32 + N/4) cycles (operand fetch; "
                             "profiles/r1_sync_probe.log): with C_out in {16,32,64} the pipe-bound ceiling of this network is ~0.65 "
                             "of dense peak",
                     "top_kernels": r["top_kernels"]},
        "elementwise": {"K1_preprocess": r.get("k1"), "K8_pack_rgb48": r.get("k8"), "hbm_peak_GBps": peaks["hbm_gbs"]},
        "clocks": r["clocks"],
        "l2": f"{r['n_distinct']} distinct input frames per rank ({r['n_distinct'] * px * 3 / 1e6:.0f} MB) cycled and a per-frame "
              "activation working set >= 1 GB: both exceed the 126 MB L2; no explicit flush",
    }
    for t in e["roofline"]["top_kernels"]:      # each launch against both ceilings: the larger fraction names its bound
        if "achieved_tflops" in t:
            t["tensor_frac_of_burst_peak"] = t["achieved_tflops"] / peaks["tflops_burst"]
        if "hbm_GBps" in t:
            t["hbm_frac_of_copy_peak"] = t["hbm_GBps"] / peaks["hbm_gbs"]
        if "tensor_frac_of_burst_peak" in t and "hbm_frac_of_copy_peak" in t:
            t["nearer_bound"] = "hbm" if t["hbm_frac_of_copy_peak"] >= t["tensor_frac_of_burst_peak"] else "tensor"
    if e["elementwise"]["K1_preprocess"]:
        for k in ("K1_preprocess", "K8_pack_rgb48"):
            e["elementwise"][k]["frac_of_hbm_peak"] = e["elementwise"][k]["achieved_GBps"] / peaks["hbm_gbs"]
    return e


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    import hdr_realtime_video_pipeline_b200 as hb
    from hdr_realtime_video_pipeline_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback for the product arm)")
    # one process per GPU: NUMA-local cores BEFORE any pinned allocation (first touch places the frame buffers)
    cores = sharding.pin_rank_to_local_cores(local_rank, local_world) if world > 1 else sorted(os.sched_getaffinity(0))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    precision = args.precision
    weights = W_INT8.get(precision, WEIGHTS)
    K, Wm = args.steps, max(3, args.warmup)
    wls = ["1080p", "4k"] if args.workload == "both" else [args.workload]
    if world > 1:
        wls = ["4k"] if args.workload == "both" else [args.workload]
    head = "4k" if "4k" in wls else wls[-1]
    peaks = measured_peaks()
    sampler = ClockSampler(local_rank)
    first, _ = hb.frame_chunk(world * K, rank, world)            # contiguous chunk of the (weak-scaled) clip
    entries, raw = {}, {}
    export = None
    with contextlib.redirect_stdout(sys.stderr):     # the wrapper prints the reference's load banner: stdout carries the JSON line only
        for wl in wls:
            net = hb.HDRTVNetB200(weights, device=f"cuda:{local_rank}", precision=precision, warmup_passes=0, use_hg=False)
            packer = hb.RGB48Packer(dev, ring_frames=3)
            r = measure_workload(net, packer, hb, torch, dev, barrier, wl, K, Wm, rank, local_rank, first, precision, sampler)
            if wl == head and not args.no_export:
                export = run_export(net, hb, sharding, torch, dist, dev, barrier, wl, K, rank, world, r["host_np"], r["n_distinct"])
            raw[wl] = r
            r.pop("host_np")
            packer.close()
            net.close()
            del net, packer
            torch.cuda.empty_cache()

    # ---- max over ranks ------------------------------------------------------------------------------
    for wl in wls:
        r = raw[wl]
        t = torch.tensor([r["dev_ms"], r["e2e_s"], r["infer_ms"], r["e2e3_s"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        r["dev_ms"], r["e2e_s"], r["infer_ms"], r["e2e3_s"] = (float(v) for v in t.tolist())
    record = {"rank": rank, "first_frame": first, "n_frames": K, "elapsed_s": raw[head]["dev_ms"] / 1000.0, "cores": cores,
              "descriptors": [(first + i, c) for i, c in enumerate(raw[head]["checks"])]}
    records = hb.gather_run_records(record)

    if rank == 0:
        for wl in wls:
            entries[wl] = workload_entry(wl, raw[wl], world, K, peaks, precision)
        hd = entries[head]
        h, w = WORKLOADS[head]
        dtype = {"fp16": "f16", "fp32": "f32", "int8-mixed": "u8 x s8 -> s32 (W8A8 convs, tcgen05.mma.kind::i8) + f16",
                 "int8-full": "f32 (fake-quantised, CUDA cores)"}[precision]
        line = {
            "metric": "HDRTVNet++ frames/sec", "value": hd["value"], "unit": "frames/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": hd["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": {"workload": CONFIG_NAME["4k_export" if (world > 1 and head == "4k") else head], "height": h, "width": w,
                       "precision": precision,
                       "weights": "HR.pt (fixture copy)" if precision in ("fp16", "fp32") else os.path.basename(W_INT8[precision]),
                       "headline": "the top-level value / e2e / roofline are the 3840x2160 workload (north-star target >= 60 frames/s); "
                                   "workloads['1080p'] carries BASELINE configs[1] incl. its frame latency",
                       "frames": "4 synthetic content classes cycled (noise / ramps / black / white+salt)",
                       "l2": hd["l2"], "sharding": f"contiguous frame chunks, {world} rank(s), no data-path collective"},
            "pixels_per_s": hd["pixels_per_s"],
            "e2e": hd["e2e"], "latency_ms": (entries.get("1080p") or hd)["latency_ms"],
            "latency_workload": "1080p" if "1080p" in entries else head,
            "gpu_launches": hd["gpu_launches"], "roofline": hd["roofline"], "clocks": hd["clocks"],
            "workloads": entries,
            "ranks": [{"rank": r["rank"], "first_frame": r["first_frame"], "n_frames": r["n_frames"], "cores": r.get("cores")}
                      for r in records],
            "host": {"cpu_count": os.cpu_count(), "cores_of_rank0": cores},
        }
        if export is not None:
            line["export"] = export
            # the end-to-end number of the line is the export loop at EVERY N (config 3 at N = 1, config 4 at N > 1), so
            # that the driver's per-N values are the same measurement
            line["e2e"] = {"value": export["frames_per_s"], "unit": "frames/s", "h2d_bytes_per_step": h * w * 3,
                           "d2h_bytes_per_step": h * w * 6 + 8,
                           "api": "hdr_realtime_video_pipeline_b200.export_clip (process_rgb48 per frame, device-side descriptor "
                                  "checksums, pinned-ring sink), one rank per GPU, contiguous frame chunks",
                           "process_rgb48_loop": hd["e2e"]["value"], "three_call_api": hd["e2e"]["three_call_api"],
                           "timing": "wall clock between device synchronisations, max over ranks"}
        if world == 1 and not args.no_gpu_baseline and precision == "fp16":
            line["gpu_eager_baseline"] = {}
            for wl in wls:
                hh, ww = WORKLOADS[wl]
                frames = [hb.synth_frame(i, hh, ww) for i in range(8)]
                line["gpu_eager_baseline"][wl] = gpu_eager_baseline(dev, hh, ww, 24 if wl == "4k" else 48, frames)
                entries[wl]["gpu_eager_baseline"] = line["gpu_eager_baseline"][wl]
        if world == 1 and precision == "fp16" and not args.no_hg:
            line["hg"] = {}
            with contextlib.redirect_stdout(sys.stderr):
                for wl in wls:
                    try:
                        line["hg"][wl] = measure_hg(hb, torch, dev, wl, min(K, 40 if wl != "4k" else 20), Wm, peaks, weights,
                                                    not args.no_gpu_baseline, with_cpu=(not args.no_cpu_baseline and wl == wls[-1]))
                    except Exception as exc:                    # the extra workload must never take the headline line down
                        line["hg"][wl] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        if world == 1 and not args.no_cpu_baseline:
            times, ncores, cpu = cpu_reference_sample(540, 960)
            cpu_px_s = 540 * 960 / float(np.mean(times))
            line["cpu_baseline"] = {
                "value": cpu_px_s / (h * w), "unit": "frames/s", "cores": ncores, "host_cpu_count": os.cpu_count(), "kind": cpu.kind,
                "pixels_per_s": cpu_px_s,
                "sample": f"{len(times)} synthetic 960x540 frames (BASELINE configs[0]) through {cpu.what}, first frame discarded, "
                          f"{float(np.mean(times)) * 1000:.0f} ms/frame; frames/s scaled by pixel count to {w}x{h}"}
        print(json.dumps(line), file=JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_export(net, hb, sharding, torch, dist, dev, barrier, wl, K, rank, world, host_np, n_distinct):
    """BASELINE configs[3]: every rank exports its contiguous chunk of a (world x K)-frame clip through export_clip (sink:
    the pinned RGB48 ring).  Frame i of the clip is synthetic frame (i mod n_distinct) of rank 0's set on EVERY rank, so the
    checksum list of a serial single-GPU pass over those frames is the oracle for the merged descriptors."""
    h, w = WORKLOADS[wl]
    from hdr_realtime_video_pipeline_b200.synth import synth_frame
    if rank == 0:
        clip = host_np
    else:                                         # same clip on every rank: frame i = synth_frame(i mod n_distinct)
        clip = [torch.from_numpy(synth_frame(i, h, w)).pin_memory().numpy() for i in range(n_distinct)]
    # oracle: serial pass on this rank's GPU, device synchronised per frame, checksum computed on the HOST from the bytes
    expected = []
    for i in range(n_distinct):
        fr = net.process_rgb48(clip[i], serial=True)
        expected.append(sharding.frame_checksum(fr.numpy()))
        fr.release()
    torch.cuda.synchronize(dev)
    n_frames = world * K
    hb.export_clip(net, lambda i: clip[i % n_distinct], min(n_frames, 4 * world), None, rank=rank, world_size=world)   # warm-up
    barrier()
    t0 = time.perf_counter()
    rec = hb.export_clip(net, lambda i: clip[i % n_distinct], n_frames, None, rank=rank, world_size=world, in_flight=3)
    torch.cuda.synchronize(dev)
    dt_local = time.perf_counter() - t0
    barrier()
    t = torch.tensor([dt_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    bad_local = [i for i, c in rec["descriptors"] if c != expected[i % n_distinct]]
    rec["bad"] = len(bad_local)
    records = hb.gather_run_records(rec)
    if rank != 0:
        return None
    merged = sharding.merge_descriptors(records)                  # raises if the chunks do not tile the clip in order
    ok = len(merged) == n_frames and all(c == expected[i % n_distinct] for i, c in merged) and not any(r["bad"] for r in records)
    if not ok:
        raise RuntimeError("sharded export: merged descriptors differ from the serial single-GPU checksum list")
    return {"config": CONFIG_NAME["4k_export"], "frames": n_frames, "frames_per_rank": K, "frames_per_s": n_frames / dt,
            "elapsed_s": dt, "ranks": world, "in_flight": 3,
            "descriptors": {"count": len(merged), "ordered": True, "match_serial_single_gpu_checksums": True,
                            "checksum": "sum_i code[i] * ((i mod 65521) + 1) over the uint16 HxWx3 frame, computed by the pack "
                                        "kernel on the GPU; the oracle list is computed on the host from the bytes of a serial pass"},
            "per_rank_elapsed_s": [r["elapsed_s"] for r in records],
            "sink": "pinned RGB48 ring (frames are released after the descriptor is taken; no disk in the loop)"}


JSON_OUT = sys.stdout


def main():
    # stdout carries the JSON line and nothing else: library banners written to file descriptor 1 (NCCL's version line,
    # the wrapper's load banner) are sent to stderr
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", choices=sorted(WORKLOADS) + ["both"], default="both")
    ap.add_argument("--precision", choices=["fp16", "fp32", "int8-mixed", "int8-full"], default="fp16")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-export", action="store_true", help="skip the export_clip leg (config 4 / the e2e figure of the line)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-hg", action="store_true", help="skip the AGCM + LE + HG workloads (line['hg'])")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
