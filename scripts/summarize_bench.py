#!/usr/bin/env python
"""Copies the bench lines a gpurun call brought back into profiles/ (tracked) with a short readable digest.
usage: python scripts/summarize_bench.py <round-tag> <ours.json> <reference.json> [<ours_2gpu.json> ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
lines = [json.loads(open(p).read().strip().splitlines()[-1]) for p in sys.argv[2:]]
out = os.path.join(ROOT, "profiles")
with open(os.path.join(out, f"{tag}_bench.jsonl"), "w") as f:
    for d in lines:
        f.write(json.dumps(d) + "\n")
md = [f"# bench.py lines, {tag}", "", "Raw JSON lines: `profiles/%s_bench.jsonl` (one per run; numbers measured by `gpurun` on a B200, not under a profiler)." % tag, "",
      "| run | n_gpus | value (evals/s) | ms/step | e2e (evals/s) | e2e ms/step | launches | SM MHz | reasons |", "|---|---:|---:|---:|---:|---:|---:|---:|---|"]
for d in lines:
    e = d.get("e2e", {})
    c = d.get("clocks", {}) or {}
    md.append(f"| {d.get('impl', 'ours')} | {d['n_gpus']} | {d['value']:.4g} | {d['ms_per_step']:.4g} | {e.get('value', 0):.4g} | "
              f"{e.get('ms_per_step', float('nan')):.4g} | {d.get('gpu_launches', '-')} | {c.get('sm_mhz', '-')} | {c.get('reasons', '-')} |")
for d in lines:
    if "roofline" in d:
        r = d["roofline"]
        md += ["", f"## roofline ({d['n_gpus']} GPU)", "", "```", json.dumps(r, indent=1), "```",
               "", "breakdown (separate profiled pass, stages of different slots overlap):", "", "```", json.dumps(d.get("breakdown_ms_per_step"), indent=1), "```",
               "", "host issue per step (ms):", "", "```", json.dumps(d.get("host_issue_ms"), indent=1), "```"]
    if "parity" in d:
        md += ["", f"## parity block of the run ({d['n_gpus']} GPU): product vs oracle on the benchmarked configuration", "", "```", json.dumps(d["parity"], indent=1), "```"]
    for k, v in (d.get("extras") or {}).items():
        md += ["", f"## extras.{k}", "", "```", json.dumps(v, indent=1), "```"]
    if "timing" in d:
        md += ["", "timing:", "", "```", json.dumps(d["timing"], indent=1), "```"]
    if "cpu_baseline" in d and d.get("impl") != "reference":
        md += ["", "cpu_baseline:", "", "```", json.dumps(d["cpu_baseline"], indent=1), "```"]
open(os.path.join(out, f"{tag}_bench.md"), "w").write("\n".join(md) + "\n")
print("wrote", f"profiles/{tag}_bench.md")
