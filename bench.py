#!/usr/bin/env python
"""bench.py — HDRTVNet++ per-frame SDR->HDR inference throughput on B200 (driver contract in the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 1080p|4k|540p] [--impl b200|reference]

A "step" is one frame through the hot path: uint8 BGR24 -> normalise -> AGCM -> LE -> RGB48 pack.
  value : frames/s with the input frames already resident in HBM (device u8 in, device u16 out), CUDA-event timed.
  e2e   : frames/s through the public API with HOST buffers: pinned u8 frame in (H2D inside the timed region),
          HDRTVNetB200.preprocess/infer, tensor_to_rgb48_bytes into the pinned ring (D2H inside the timed region).
  N > 1 : one process per GPU (torchrun), each rank runs the same K steps on its own contiguous frame chunk (weak
          scaling, frames are independent — no collective on the pixel path); NCCL barrier on both sides, MAX over ranks.
  --impl reference : the reference's CPU eager path restated in oracle/torch_port.py (the reference tree itself is Python
          and does not exist on the GPU box), all host threads, bounded sample per step.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {"540p": (540, 960), "1080p": (1080, 1920), "4k": (2160, 3840)}
CONFIG_NAME = {
    "540p": "HDRTVNet++ (AGCM+LE) 960x540 synthetic BT.709 SDR frames (BASELINE configs[0] size)",
    "1080p": "HDRTVNet++ FP16 1920x1080 single-B200 real-time playback stream (batch 1) — BASELINE configs[1]",
    "4k": "HDRTVNet++ FP16 3840x2160 single-B200 inference with fused RGB48 pack into pinned host ring — BASELINE configs[2]",
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
# dram__bytes_read.sum + dram__bytes_write.sum per launch at 1920x1080 from profiles/r1_ncu_top_kernels.md (ncu --set full)
NCU_TRAFFIC_1080P = {}     # filled from profiles/r1_ncu_top_kernels.json when present (written by scripts/summarise_ncu.py)
NCU_FRAME_TRAFFIC = {}     # workload -> dram read + write bytes of every launch of one frame (profiles/r1_traffic_1080p.csv, r1_launches_4k.csv)
try:
    with open(os.path.join(REPO, "profiles", "r1_ncu_top_kernels.json")) as _f:
        _j = json.load(_f)
        NCU_TRAFFIC_1080P = {k: int(v) for k, v in _j.get("dram_bytes_per_launch_1080p", {}).items()}
        for _wl in ("1080p", "4k"):
            if _j.get(f"frame_dram_bytes_{_wl}", {}).get("total"):
                NCU_FRAME_TRAFFIC[_wl] = int(_j[f"frame_dram_bytes_{_wl}"]["total"])
except Exception:
    pass
WEIGHTS = os.path.join(REPO, "tests", "golden", "weights_hr.npz")


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops": float(p.get("bf16_tflops_sustained", 1381.4)), "hbm_gbs": float(p.get("hbm_gbs", 6556.2)),
                "source": "MEASURED_PEAKS.json (cuBLAS bf16 sustained, STREAM-style copy)"}
    return {"tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for t, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8 or not (t0 <= t <= t1 + 0.2):
                continue
            try:
                sm.append(float(parts[1]))
                mx = max(mx, float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU baseline legs
def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs are meant to use every core this process may run on."""
    import torch
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def cpu_reference_sample(n_frames: int, height: int, width: int, budget_s: float = 25.0):
    """The reference's setup_cpu configuration (torch CPU eager fp32) through oracle/torch_port.py."""
    import torch
    from oracle import torch_port as TP
    _use_all_host_threads()
    from hdr_realtime_video_pipeline_b200.synth import synth_frame
    sd = TP.to_torch_state(dict(np.load(WEIGHTS)))
    TP.process_rgb48(sd, synth_frame(0, height, width))          # first frame discarded
    times, t_start = [], time.perf_counter()
    for i in range(n_frames):
        f = synth_frame(i + 1, height, width)
        t0 = time.perf_counter()
        TP.process_rgb48(sd, f)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s and len(times) >= 2:
            break
    return times, torch.get_num_threads()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    _use_all_host_threads()
    h, w = WORKLOADS[args.workload]
    sh, sw = 540, 960                                           # bounded sample: one config-1-size frame per step
    from oracle import torch_port as TP
    from hdr_realtime_video_pipeline_b200.synth import synth_frame
    sd = TP.to_torch_state(dict(np.load(WEIGHTS)))
    for i in range(max(1, args.warmup)):
        TP.process_rgb48(sd, synth_frame(i, sh, sw))
    t0 = time.perf_counter()
    for i in range(args.steps):
        TP.process_rgb48(sd, synth_frame(100 + i, sh, sw))
    dt = time.perf_counter() - t0
    px_per_s = args.steps * sh * sw / dt
    fps = px_per_s / (h * w)
    cores = torch.get_num_threads()
    sample = (f"each step = one {sw}x{sh} frame (1/{(h * w) // (sh * sw)} of a {w}x{h} frame) through the torch-CPU fp32 port of the "
              f"reference's eager path; frames/s scaled by pixel count to {w}x{h}")
    line = {"impl": "reference", "metric": "HDRTVNet++ frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1000.0, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIG_NAME[args.workload], "height": h, "width": w},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=JSON_OUT, flush=True)


# ---------------------------------------------------------------------------------------------- B200 arm
def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    import hdr_realtime_video_pipeline_b200 as hb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback for the product arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    h, w = WORKLOADS[args.workload]
    precision = "fp32" if args.precision == "fp32" else "fp16"
    with contextlib.redirect_stdout(sys.stderr):     # the wrapper prints the reference's load banner: stdout carries the JSON line only
        net = hb.HDRTVNetB200(WEIGHTS, device=f"cuda:{local_rank}", precision=precision, warmup_passes=0, use_hg=False)
    packer = hb.RGB48Packer(dev, ring_frames=3)
    K, Wm = args.steps, max(3, args.warmup)
    n_distinct = 16
    first, _ = hb.frame_chunk(world * K, rank, world)            # contiguous chunk of the (weak-scaled) clip
    host_frames = [torch.from_numpy(hb.synth_frame(first + i, h, w)).pin_memory() for i in range(n_distinct)]
    dev_frames = [f.to(dev) for f in host_frames]
    out_dev = torch.empty((h, w, 3), dtype=torch.uint16, device=dev)

    def step_device(i):
        x, c = net.preprocess_device(dev_frames[i % n_distinct], assume_ready=True)
        out = net.infer((x, c))
        packer.pack_device(out, out_dev)

    # ---- device-resident throughput -----------------------------------------------------------------
    for i in range(Wm):
        step_device(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
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
    dev_ms = ev0.elapsed_time(ev1)
    launches = net.launch_count() + packer.launch_count() - l0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    infer_ms = float(np.mean([a.elapsed_time(b) for a, b in infer_ev]))

    # ---- end to end through the public API, host buffers --------------------------------------------
    # (a) the one-call frame path (HDRTVNetB200.process_rgb48 -> hdrtv_process): pinned u8 frame in, RGB48 in a pinned
    #     ring slot out, H2D and D2H inside the timed region, two frames in flight
    # (b) the reference's three calls (preprocess -> infer -> _tensor_to_rgb48_bytes), same buffers
    state = {}

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

    host_np = [f.numpy() for f in host_frames]
    one_call_serial = os.environ.get("BENCH_ONE_CALL_SERIAL", "0") == "1"       # diagnostic knob
    e2e_s, checks = e2e_run(lambda i: net.process_rgb48(host_np[i % n_distinct], serial=one_call_serial), 3)
    e2e3_s, checks3 = e2e_run(lambda i: hb.tensor_to_rgb48_bytes(net.infer(net.preprocess(host_np[i % n_distinct])), state), 2)
    if checks != checks3:
        raise RuntimeError("one-call and three-call frame paths disagree")

    # ---- batch-1 latency (config 2 is latency-bound): serial frames, host-timed submit .. RGB48 slot ready --------
    def latency_run(submit):
        lat = []
        for i in range(min(K, 60)):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            fr = submit(i)
            fr.wait_ready()
            lat.append((time.perf_counter() - t0) * 1000.0)
            fr.release()
        return lat

    lat = latency_run(lambda i: net.process_rgb48(host_np[i % n_distinct], serial=True))
    lat3 = latency_run(lambda i: hb.tensor_to_rgb48_bytes(net.infer(net.preprocess(host_np[i % n_distinct])), state))

    # ---- per-launch device times of one frame (CUDA events between launches, median of 5) -> top-kernel rooflines
    top_kernels = []
    if rank == 0 and precision == "fp16":
        x, c = net.preprocess_device(dev_frames[0], assume_ready=True)
        runs = [net.time_plan((x, c)) for _ in range(5)]
        names = [n for n, _ in runs[0]]
        med = np.median(np.array([[t for _, t in r] for r in runs]), axis=0)
        order = np.argsort(-med)[:4]
        for i in order:
            name = names[i].split(" ")[0]
            mac = KERNEL_MAC_PER_PX.get(name)
            entry = {"launch": names[i], "ms": float(med[i]), "share_of_infer": float(med[i] / med.sum())}
            if mac is not None:
                tf = 2.0 * mac * h * w / (med[i] / 1000.0) / 1e12
                entry.update({"algorithmic_mac_per_px": mac, "achieved_tflops": tf})
            traffic = NCU_TRAFFIC_1080P.get(name)
            if traffic is not None and (h, w) == (1080, 1920):
                entry["traffic_bytes_per_launch"] = traffic
            top_kernels.append(entry)

    # ---- max over ranks ------------------------------------------------------------------------------
    t = torch.tensor([dev_ms, e2e_s * 1000.0, infer_ms, e2e3_s * 1000.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, infer_ms, e2e3_ms = (float(v) for v in t.tolist())
    record = {"rank": rank, "first_frame": first, "n_frames": K, "elapsed_s": dev_ms / 1000.0,
              "descriptors": [(first + i, c) for i, c in enumerate(checks)]}
    records = hb.gather_run_records(record)

    if rank == 0:
        peaks = measured_peaks()
        px = h * w
        fps = world * K / (dev_ms / 1000.0)
        e2e_fps = world * K / (e2e_ms / 1000.0)
        step_ms = dev_ms / K                                   # whole step: preprocess + infer + pack, device-timed
        achieved_tf = FLOP_PER_PX * px / (step_ms / 1000.0) / 1e12
        frame_traffic = NCU_FRAME_TRAFFIC.get(args.workload) if precision == "fp16" else None
        line = {
            "metric": "HDRTVNet++ frames/sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if precision == "fp16" else "f32", "data": "synthetic",
            "config": {"workload": CONFIG_NAME[args.workload], "height": h, "width": w, "weights": "HR.pt (fixture copy)",
                       "frames": "4 synthetic content classes cycled, 16 distinct frames per rank",
                       "l2": "per-frame activation working set (>= 1 GB at 1080p) exceeds the 126 MB L2; no explicit flush",
                       "sharding": f"contiguous frame chunks, {world} rank(s), no data-path collective"},
            "pixels_per_s": fps * px,
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": px * 3, "d2h_bytes_per_step": px * 6,
                    "api": "HDRTVNetB200.process_rgb48 (one C-ABI call per frame: hdrtv_process), pinned u8 frame in, RGB48 in a "
                           "pinned ring slot out",
                    "three_call_api": {"value": world * K / (e2e3_ms / 1000.0), "unit": "frames/s",
                                       "api": "preprocess -> infer -> tensor_to_rgb48_bytes (the reference's call sequence)"},
                    "timing": "wall clock between device synchronisations, includes host-side launch cost"},
            "latency_ms": {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)),
                           "three_call_api_p50": float(np.percentile(lat3, 50)),
                           "what": "pinned u8 in -> RGB48 in pinned ring slot ready, one frame in flight (process_rgb48 serial=True)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": achieved_tf / peaks["tflops"],
                         "traffic": frame_traffic,
                         "traffic_note": "DRAM read + write bytes of all launches of one frame (ncu, profiles/r1_traffic_1080p.csv / "
                                         "r1_launches_4k.md): inter-layer activations; the frame's algorithmic I/O is 9 B/px",
                         "hbm": ({"achieved_GBps": frame_traffic / (step_ms * 1e-3) / 1e9, "peak_GBps": peaks.get("hbm_gbs"),
                                  "frac": (frame_traffic / (step_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if peaks.get("hbm_gbs") else None}
                                 if frame_traffic else None),
                         "kernel": "whole hot path of one frame (chain_p8_kernel x2, conv2x_p8_kernel, conv_p8_kernel family, "
                                   "classifier, pre/pack); CUDA events around the timed steps on the launching stream",
                         "algorithmic": f"{FLOP_PER_PX:.0f} FLOP/px x {px} px per frame", "peak_source": peaks["source"],
                         "ms_per_frame": step_ms, "infer_only_ms": infer_ms,
                         "note": "tcgen05.mma M=128 K=16 occupies the pipe for max(N/2, 32 + N/4) cycles (operand fetch; "
                                 "profiles/r1_sync_probe.log): with C_out in {16,32,64} the pipe-bound ceiling of this network is ~0.65 "
                                 "of dense peak",
                         "top_kernels": top_kernels},
            "clocks": clocks,
            "ranks": [{"rank": r["rank"], "first_frame": r["first_frame"], "n_frames": r["n_frames"]} for r in records],
        }
        if world == 1 and not args.no_cpu_baseline:
            times, cores = cpu_reference_sample(8, 540, 960)
            cpu_px_s = 540 * 960 / float(np.mean(times))
            line["cpu_baseline"] = {
                "value": cpu_px_s / px, "unit": "frames/s", "cores": cores, "kind": "port",
                "sample": f"{len(times)} synthetic 960x540 frames (BASELINE configs[0]) through oracle/torch_port.py "
                          f"(torch CPU eager fp32, the reference's setup_cpu path), first frame discarded, "
                          f"{float(np.mean(times)) * 1000:.0f} ms/frame; frames/s scaled by pixel count to {w}x{h}"}
        print(json.dumps(line), file=JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="1080p")
    ap.add_argument("--precision", choices=["fp16", "fp32"], default="fp16")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
