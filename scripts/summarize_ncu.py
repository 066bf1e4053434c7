#!/usr/bin/env python
"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/.
usage: python scripts/summarize_ncu.py <round-tag> <launches.csv> <full.ncu-rep> [launch-list command] [full-capture command]"""
import csv
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
cmd_launches = sys.argv[4] if len(sys.argv) > 4 else "ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c 400 --csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline"
cmd_full = sys.argv[5] if len(sys.argv) > 5 else "ncu --set full --clock-control none --import-source on -k regex:<kernels> -s <skip> -c <n> python bench.py --steps 12 --warmup 3 --no-cpu-baseline"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list: per-kernel count / time / share ----
rows = [r for r in csv.reader(open(launches)) if len(r) > 14 and r[0].isdigit()]
agg = OrderedDict()
for r in rows:
    name = r[4].split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += float(r[14])
tot = sum(v[1] for v in agg.values())
lines = [f"# ncu launch list, {tag}", "",
         f"Command: `{cmd_launches}`",
         f"({len(rows)} launches captured after the priming/warm-up frames; per-launch times are cold-cache and serialised: compare SHARES)", "",
         "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| `{k}` | {n} | {t / 1e3:.1f} | {t / n / 1e3:.2f} | {100 * t / tot:.1f} % |")
open(os.path.join(out_dir, f"{tag}_launches.md"), "w").write("\n".join(lines) + "\n")
with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w") as f:
    f.write(open(launches).read())

# ---- full capture: key metrics per profiled kernel ----
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr = rr[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_active.max",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__occupancy_limit_registers"]
idx = {h: i for i, h in enumerate(hdr)}
units = rr[1]
lines = [f"# ncu --set full, {tag}", "",
         f"Command: `{cmd_full}`",
         "", "traffic = dram__bytes_read.sum + dram__bytes_write.sum per launch.", ""]
for r in rr[2:]:
    if len(r) < len(hdr):
        continue
    lines.append(f"## {r[idx['Kernel Name']].split('(')[0]}")
    lines.append("")
    lines.append("| metric | value | unit |")
    lines.append("|---|---:|---|")
    for w in want[1:]:
        if w in idx:
            lines.append(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |")
    stalls = [(h, float(r[i])) for h, i in idx.items() if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and r[i].replace('.', '', 1).isdigit()]
    tot_s = sum(v for _, v in stalls) or 1.0
    top = sorted(stalls, key=lambda kv: -kv[1])[:5]
    lines.append("")
    lines.append("top stall reasons (pc samples): " + ", ".join(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * v / tot_s:.0f} %" for h, v in top))
    lines.append("")
open(os.path.join(out_dir, f"{tag}_full_summary.md"), "w").write("\n".join(lines) + "\n")
print("wrote", [f for f in os.listdir(out_dir) if f.startswith(tag)])
