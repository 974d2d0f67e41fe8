#!/usr/bin/env python
"""Print the headline numbers of a bench.py JSON line (one per workload).   python scripts/show_bench.py line.json"""
import json
import sys

d = json.load(open(sys.argv[1]))
print(f"n_gpus {d['n_gpus']} value {d['value']:.1f} e2e {d['e2e']['value']:.1f} {d['unit']}  ({d['config']['workload'][:60]})")
for wl, e in d.get("workloads", {}).items():
    lat = e["latency_ms"]
    ck = e.get("clocks") or {}
    print(f"{wl}: value {e['value']:.1f} e2e {e['e2e']['value']:.1f} three-call {e['e2e']['three_call_api']['value']:.1f} ms/step {e['ms_per_step']:.3f} "
          f"frac {e['roofline']['frac']:.3f} p50 {lat['p50']:.2f} p99 {lat['p99']:.2f} (3-call p50 {lat['three_call_api_p50']:.2f}) "
          f"clk {ck.get('sm_mhz')} {ck.get('reasons')}")
    el = e.get("elementwise") or {}
    if el.get("K1_preprocess"):
        print(f"   K1 {el['K1_preprocess']['achieved_GBps']:.0f} GB/s  K8 {el['K8_pack_rgb48']['achieved_GBps']:.0f} GB/s")
    for t in e["roofline"]["top_kernels"]:
        print(f"   {t['launch'][:44]:44s} {t['ms']:.4f} ms {t.get('achieved_tflops', 0):.0f} TF")
    if e.get("gpu_eager_baseline"):
        print("   gpu eager reference:", {k: v for k, v in e["gpu_eager_baseline"].items() if k != "what"})
if d.get("export"):
    print("export:", d["export"]["frames"], "frames", f"{d['export']['frames_per_s']:.1f} frames/s", d["export"]["descriptors"]["match_serial_single_gpu_checksums"])
if d.get("cpu_baseline"):
    print("cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"])
for wl, e in (d.get("hg") or {}).items():
    if "error" in e:
        print(f"hg {wl}: ERROR {e['error']}")
        continue
    rf = e["roofline"]
    print(f"hg {wl}: value {e['value']:.1f} e2e {e['e2e']['value']:.1f} frames/s ms/step {e['ms_per_step']:.3f}; HG stage {rf['ms_per_frame']:.3f} ms "
          f"{rf['achieved']:.0f} TFLOP/s frac {rf['frac']:.3f} (burst {rf['frac_of_burst_peak']:.3f})")
    for t in rf["top_kernels"]:
        print(f"   {t['launch'][:44]:44s} {t['ms']:.4f} ms {t['achieved_tflops']:.0f} TF")
    if e.get("mask_early_out"):
        print(f"   with the highlight-mask early-out (product default, not the headline): {e['mask_early_out']['value']:.1f} frames/s")
    if e.get("cpu_baseline"):
        print("   cpu reference (with HG):", {k: v for k, v in e["cpu_baseline"].items() if k not in ("sample",)})
    if e.get("gpu_eager_baseline"):
        print("   gpu eager reference (with HG):", {k: v for k, v in e["gpu_eager_baseline"].items() if k != "what"})
