#!/usr/bin/env python
"""Launch list + DRAM traffic of one frame from an ncu CSV with several metrics per launch.

    python scripts/summarise_launches.py gpurun_out/launches_4k.csv 4k [frames]

The CSV comes from
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file <csv> python scripts/profile_frame.py 4k 1
(3 warm-up frames + 1 profiled frame = 4 frames).  Writes profiles/r1_launches_<tag>.{csv,md} and merges
`frame_dram_bytes_<tag>` into profiles/r1_ncu_top_kernels.json (read by bench.py for roofline.traffic)."""
import collections
import csv
import json
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(REPO, "profiles")
src, tag = sys.argv[1], sys.argv[2]
nframes = int(sys.argv[3]) if len(sys.argv) > 3 else 4

rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = rows[0]
iI, iK, iM, iU, iV = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
iG, iB = hdr.index("Grid Size"), hdr.index("Block Size")
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0,
         "msecond": 1e3}
launch = collections.OrderedDict()
for r in rows[1:]:
    d = launch.setdefault(int(r[iI]), {"kernel": r[iK].split("(")[0].replace("void ", ""), "grid": r[iG], "block": r[iB]})
    d[r[iM]] = float(r[iV].replace(",", "")) * SCALE.get(r[iU], 1.0)
data = list(launch.values())
per = len(data) // nframes
last = data[-per:]
t_tot = sum(d["gpu__time_duration.sum"] for d in last)
rd = sum(d.get("dram__bytes_read.sum", 0.0) for d in last)
wr = sum(d.get("dram__bytes_write.sum", 0.0) for d in last)

shutil.copy(src, os.path.join(PROF, f"r1_launches_{tag}.csv"))
agg = collections.OrderedDict()
for d in last:
    a = agg.setdefault(d["kernel"], [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d["gpu__time_duration.sum"]
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
with open(os.path.join(PROF, f"r1_launches_{tag}.md"), "w") as f:
    f.write(f"# Round 1 — launch list and DRAM traffic of one {tag} FP16 frame\n\n")
    f.write("Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv "
            f"--log-file gpurun_out/launches_{tag}.csv python scripts/profile_frame.py {tag} 1`\n")
    f.write(f"(raw CSV: r1_launches_{tag}.csv, {nframes} frames x {per} launches; tables = the last frame).  Times under ncu are "
            "cold-cache and serialised: shares matter, not absolutes.\n\n")
    f.write(f"Launches per frame: {per}; sum of kernel durations {t_tot:.1f} us; DRAM traffic of the frame {(rd + wr) / 1e9:.2f} GB "
            f"({rd / 1e9:.2f} GB read, {wr / 1e9:.2f} GB written).\n\n")
    f.write("| share | us | launches | DRAM MB | GB/s | kernel |\n|---|---|---|---|---|---|\n")
    for k, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {100 * t / t_tot:.1f}% | {t:.1f} | {n} | {b / 1e6:.0f} | {b / t / 1e3:.0f} | `{k}` |\n")
    f.write("\n## Per-launch list (last frame)\n\n| # | us | DRAM read MB | DRAM write MB | GB/s | grid | block | kernel |\n|---|---|---|---|---|---|---|---|\n")
    for i, d in enumerate(last):
        r_, w_ = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
        t = d["gpu__time_duration.sum"]
        f.write(f"| {i} | {t:.1f} | {r_ / 1e6:.1f} | {w_ / 1e6:.1f} | {(r_ + w_) / t / 1e3:.0f} | {d['grid']} | {d['block']} | `{d['kernel']}` |\n")

jp = os.path.join(PROF, "r1_ncu_top_kernels.json")
j = json.load(open(jp)) if os.path.isfile(jp) else {}
j[f"frame_dram_bytes_{tag}"] = {"read": int(rd), "write": int(wr), "total": int(rd + wr), "launches": per,
                                "sum_kernel_us_under_ncu": t_tot}
json.dump(j, open(jp, "w"), indent=1)
print(f"{tag}: {per} launches/frame, {t_tot:.1f} us under ncu, DRAM {(rd + wr) / 1e9:.2f} GB")
