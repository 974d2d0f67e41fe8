#!/usr/bin/env python
"""Key metrics of `ncu --set full` captures taken at 3840x2160 -> profiles/r1_ncu_top_kernels_4k.md
    python scripts/summarise_ncu_4k.py gpurun_out/a.ncu-rep [b.ncu-rep ...]"""
import csv
import io
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct_of_peak"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_read_bytes"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct_of_peak"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "tensor_cycles_active_per_sm"),
        ("sm__cycles_elapsed.max", "cycles_elapsed"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct")]
md = ["# Round 1 — ncu full captures of the top kernels at 3840x2160 (key metrics)\n",
      "Command: `ncu --set full --clock-control none --import-source on -k <kernel> -s <skip> -c <n> python scripts/profile_frame.py 4k 1`."
      "  Per-launch times under ncu are cold-cache and serialised.  `tensor_cycles_active_per_sm` sums the four sub-partition pipes:"
      " an M=128 x N=64 x K=16 MMA counts 128 (= 4 x 32 ideal cycles) while it occupies the pipe for 48, so pipe-busy = active / 4 x 48/32 / elapsed"
      " for N = 64 kernels.\n"]
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rws = list(csv.reader(io.StringIO(txt)))
    if len(rws) < 3:
        continue
    h, units = rws[0], rws[1]
    md.append(f"\n## {os.path.basename(rep)}\n")
    for r in rws[2:]:
        md.append(f"\n**{r[h.index('Kernel Name')]}**\n")
        vals = {}
        for key, label in WANT:
            cand = [i for i, col in enumerate(h) if col == key] + [i for i, col in enumerate(h) if col != key and col.endswith(key)]
            for i in cand:
                if r[i].strip() not in ("", "no data"):
                    vals[label] = r[i]
                    md.append(f"- {label}: {r[i]} {units[i]}")
                    break
        try:
            ta = float(vals["tensor_cycles_active_per_sm"].replace(",", ""))
            ce = float(vals["cycles_elapsed"].replace(",", ""))
            md.append(f"- tensor math cycles / (4 x elapsed): {100 * ta / ce / 4:.1f} % of the dense rate")
        except Exception:
            pass
open(os.path.join(REPO, "profiles", "r1_ncu_top_kernels_4k.md"), "w").write("\n".join(md) + "\n")
