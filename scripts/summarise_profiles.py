#!/usr/bin/env python
"""Summaries of ncu output for profiles/ (run in the build container after a gpurun profiling call).

    python scripts/summarise_profiles.py launches <round-tag> <workload> <launches.csv> [frames-in-csv]
    python scripts/summarise_profiles.py full     <round-tag> <workload> <raw-page.csv | report.ncu-rep> [...]
    python scripts/summarise_profiles.py sass     <round-tag>

launches: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv
          --log-file <csv> python scripts/profile_frame.py <workload> 1      (3 warm-up frames + 1 profiled frame)
          -> profiles/<tag>_launches_<workload>.{csv,md}, whole-frame and per-launch DRAM bytes merged into
             profiles/<tag>_traffic.json (read by bench.py for roofline.traffic)
full:     ncu --set full --clock-control none ... ; either the .ncu-rep or its `--page raw --csv` export
          -> profiles/<tag>_ncu_<workload>.md.  Tensor-pipe figures use metrics that cannot exceed 100 %:
             sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed   math cycles of the dense fp16 rate (8192 FLOP/clk/SM)
             sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed       tensor-core pipe busy (math + operand fetch)
             l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_...     shared-memory operand bandwidth used by the MMAs
sass:     cuobjdump -sass opcode histogram of the product library -> profiles/<tag>_sass_histogram.md
"""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(REPO, "profiles")
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0,
         "msecond": 1e3, "second": 1e6}


def short(k):
    return k.split("(")[0].replace("void ", "").replace("hdrtv::", "")


def launch_names(workload):
    """Launch names of one fp16 frame in order, from a time_layers log of the same build if one is around (optional)."""
    for cand in sorted((f for f in os.listdir(os.path.join(REPO, "gpurun_out")) if "layers" in f and workload in f and "int8" not in f),
                       key=lambda f: -os.path.getmtime(os.path.join(REPO, "gpurun_out", f))):
        names = [ln.split(" us  ", 1)[1].strip() for ln in open(os.path.join(REPO, "gpurun_out", cand)) if " us  " in ln]
        if names:
            return names, cand
    return None, None


def do_launches(tag, wl, src, nframes=4):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    iI, iK, iM, iU, iV = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    iG, iB = hdr.index("Grid Size"), hdr.index("Block Size")
    launch = collections.OrderedDict()
    for r in rows[1:]:
        d = launch.setdefault(int(r[iI]), {"kernel": short(r[iK]), "grid": r[iG], "block": r[iB]})
        d[r[iM]] = float(r[iV].replace(",", "")) * SCALE.get(r[iU], 1.0)
    data = list(launch.values())
    per = len(data) // nframes
    last = data[-per:]
    t_tot = sum(d["gpu__time_duration.sum"] for d in last)
    rd = sum(d.get("dram__bytes_read.sum", 0.0) for d in last)
    wr = sum(d.get("dram__bytes_write.sum", 0.0) for d in last)
    shutil.copy(src, os.path.join(PROF, f"{tag}_launches_{wl}.csv"))
    agg = collections.OrderedDict()
    for d in last:
        a = agg.setdefault(d["kernel"], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d["gpu__time_duration.sum"]
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    # the launch plan's own names (front end: normalise+P8, 5 classifier levels, head; then the plan; then the pack)
    names, name_src = launch_names(wl)
    with open(os.path.join(PROF, f"{tag}_launches_{wl}.md"), "w") as f:
        f.write(f"# {tag} — launch list and DRAM traffic of one {wl} FP16 frame\n\n")
        f.write("Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv "
                f"--log-file gpurun_out/<csv> python scripts/profile_frame.py {wl} 1`\n")
        f.write(f"(raw CSV: {tag}_launches_{wl}.csv, {nframes} frames x {per} launches; tables = the last frame).  Times under ncu are "
                "cold-cache and serialised: shares matter, not absolutes.\n\n")
        f.write(f"Launches per frame: {per}; sum of kernel durations {t_tot:.1f} us; DRAM traffic of the frame {(rd + wr) / 1e9:.2f} GB "
                f"({rd / 1e9:.2f} GB read, {wr / 1e9:.2f} GB written).\n\n")
        f.write("| share | us | launches | DRAM MB | GB/s | kernel |\n|---|---|---|---|---|---|\n")
        for k, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {100 * t / t_tot:.1f}% | {t:.1f} | {n} | {b / 1e6:.0f} | {b / t / 1e3:.0f} | `{k}` |\n")
        f.write("\n## Per-launch list (last frame)\n\n| # | us | share | DRAM read MB | DRAM write MB | GB/s | grid | block | kernel |\n"
                "|---|---|---|---|---|---|---|---|---|\n")
        for i, d in enumerate(last):
            r_, w_ = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
            t = d["gpu__time_duration.sum"]
            f.write(f"| {i} | {t:.1f} | {100 * t / t_tot:.1f}% | {r_ / 1e6:.1f} | {w_ / 1e6:.1f} | {(r_ + w_) / t / 1e3:.0f} | {d['grid']} | "
                    f"{d['block']} | `{d['kernel']}` |\n")
    # per-launch traffic keyed by the plan's launch names: the plan's tensor-core launches are the chain / conv / conv2x kernels in order
    named = {}
    if names:
        plan = [n for n in names if n.split(" ")[0] not in ("planar_to_p8", "agcm_head") and not n.startswith("cls.")]
        tens = [d for d in last if d["kernel"].startswith(("chain_p8_kernel", "conv_p8_kernel", "conv2x_p8_kernel", "conv3z_pair_kernel"))]
        if len(plan) == len(tens):
            for n, d in zip(plan, tens):
                named[n.split(" ")[0]] = int(d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0))
    jp = os.path.join(PROF, f"{tag}_traffic.json")
    j = json.load(open(jp)) if os.path.isfile(jp) else {}
    j[f"frame_dram_bytes_{wl}"] = {"read": int(rd), "write": int(wr), "total": int(rd + wr), "launches": per,
                                   "sum_kernel_us_under_ncu": t_tot}
    j[f"dram_bytes_per_launch_{wl}"] = named
    json.dump(j, open(jp, "w"), indent=1)
    print(f"{wl}: {per} launches/frame, {t_tot:.1f} us under ncu, DRAM {(rd + wr) / 1e9:.2f} GB; named launches: {len(named)} (names from {name_src})")


WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct_of_peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_math_pct_of_dense_rate"),
        ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_core_pipe_busy_pct"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "mma_smem_operand_bandwidth_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct_of_peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct")]


def raw_rows(path):
    if path.endswith(".ncu-rep"):
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        txt = open(path).read()
    rws = list(csv.reader(io.StringIO(txt)))
    return rws if len(rws) >= 3 else None


def do_full(tag, wl, paths):
    md = [f"# {tag} — `ncu --set full` key metrics, {wl} FP16 frame\n",
          "Command: `ncu --set full --clock-control none [-k regex:<kernel> -s <skip> -c <n>] python scripts/profile_frame.py "
          f"{wl} 1` (captured after the same command had exited 0 without ncu).  Per-launch times under ncu are cold-cache and "
          "serialised.\n",
          "Tensor-pipe columns (none of them can exceed 100 %): **math** = `sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed` "
          "(cycles the MMA datapath computes, as a share of elapsed: equals achieved / dense fp16 rate of 8192 FLOP/clk/SM); **pipe busy** = "
          "`sm__pipe_tc_cycles_active...` (the tensor-core pipe incl. operand fetch: an M=128 N=64 K=16 MMA holds it for 48 cycles and "
          "computes for 32); **smem operand bw** = `l1tex__data_pipe_tc_wavefronts_mem_shared...` (shared-memory bandwidth drawn by MMA "
          "operands).\n"]
    table = ["| kernel | grid | us | DRAM MB (r+w) | DRAM % | L2 % | math % | pipe busy % | smem operand bw % | regs |", "|---|---|---|---|---|---|---|---|---|---|"]
    for path in paths:
        rws = raw_rows(path)
        if not rws:
            continue
        h, units = rws[0], rws[1]
        idx = {}
        for key, label in WANT:
            for i, col in enumerate(h):
                if col == key:
                    idx[label] = i
                    break
        for r in rws[2:]:
            def val(label):
                i = idx.get(label)
                if i is None or r[i].strip() in ("", "no data"):
                    return None
                return float(r[i].replace(",", "")) * (SCALE.get(units[i], 1.0) if label in ("duration", "dram_read", "dram_write") else 1.0)
            name = short(r[h.index("Kernel Name")])
            rd, wr = val("dram_read") or 0.0, val("dram_write") or 0.0

            def fmt(label, nd=1):
                x = val(label)
                return "-" if x is None else f"{x:.{nd}f}"
            table.append(f"| `{name[:90]}` | {fmt('grid', 0)} | {fmt('duration')} | {(rd + wr) / 1e6:.0f} | {fmt('dram_pct_of_peak')} | "
                         f"{fmt('l2_pct_of_peak')} | {fmt('tensor_math_pct_of_dense_rate')} | {fmt('tensor_core_pipe_busy_pct')} | "
                         f"{fmt('mma_smem_operand_bandwidth_pct')} | {fmt('regs', 0)} |")
    md += table
    out = os.path.join(PROF, f"{tag}_ncu_{wl}.md")
    open(out, "w").write("\n".join(md) + "\n")
    print("wrote", out, len(table) - 2, "kernels")


def do_sass(tag):
    lib = os.path.join(REPO, "hdr_realtime_video_pipeline_b200", "libhdrtv_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    ops = collections.Counter()
    fn = collections.Counter()
    cur = None
    per_fn = collections.defaultdict(collections.Counter)
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m:
            op = m.group(1)
            ops[op] += 1
            per_fn[cur][op] += 1
    keys = ["UTCHMMA", "UTCIMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "HMMA", "IMMA", "HGMMA", "SYNCS", "LDGSTS"]
    md = [f"# {tag} — SASS opcode histogram of libhdrtv_b200.so (sm_100a)\n",
          "Command: `cuobjdump -sass hdr_realtime_video_pipeline_b200/libhdrtv_b200.so` (scripts/summarise_profiles.py sass).  "
          "`UTCHMMA` / `UTCIMMA` = tcgen05.mma kind::f16 / kind::i8, `LDTM` / `STTM` = tcgen05.ld / st, `UBLKCP` = cp.async.bulk (1-D bulk "
          "TMA), `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier; `HMMA` / `IMMA` (legacy mma.sync) must be absent.\n",
          "| opcode | count |", "|---|---|"]
    for k in keys:
        md.append(f"| {k} | {ops.get(k, 0)} |")
    md.append(f"\nFunctions: {len(per_fn)}; instructions: {sum(ops.values())}.\n")
    md.append("| kernel family | instances | UTCHMMA | UTCIMMA | LDTM | STTM | UBLKCP |\n|---|---|---|---|---|---|---|")
    fam = collections.defaultdict(lambda: [0, collections.Counter()])
    for name, c in per_fn.items():
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() if name else ""
        base = re.sub(r"<.*", "", short(dem.replace("hdrtv::", ""))) or name
        fam[base][0] += 1
        fam[base][1].update(c)
    for base, (n, c) in sorted(fam.items(), key=lambda kv: -kv[1][1].get("UTCHMMA", 0) - kv[1][1].get("UTCIMMA", 0)):
        md.append(f"| `{base}` | {n} | {c.get('UTCHMMA', 0)} | {c.get('UTCIMMA', 0)} | {c.get('LDTM', 0)} | {c.get('STTM', 0)} | {c.get('UBLKCP', 0)} |")
    md.append("\nTop 25 opcodes: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(25)) + "\n")
    out = os.path.join(PROF, f"{tag}_sass_histogram.md")
    open(out, "w").write("\n".join(md) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    mode, tag = sys.argv[1], sys.argv[2]
    if mode == "launches":
        do_launches(tag, sys.argv[3], sys.argv[4], int(sys.argv[5]) if len(sys.argv) > 5 else 4)
    elif mode == "full":
        do_full(tag, sys.argv[3], sys.argv[4:])
    elif mode == "sass":
        do_sass(tag)
    else:
        raise SystemExit(__doc__)
