#!/usr/bin/env python
"""profiles/r2_lm_counters.json from ncu captures: per-launch FP64 operation counts and DRAM bytes of the dominant kernel
(lm2_kernel), which bench.py combines with the evaluation counter and the isolated launch time it measures itself.
usage: python scripts/ncu_lm_counters.py <one-frame.ncu-rep> <saturated.ncu-rep> <evals executed in the one-frame launch> <evals in the saturated launch>"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows(rep, kre):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    return rr[0], rr[2:]


def counters(rep, kre):
    h, rs = rows(rep, kre)
    r = rs[0]
    g = lambda name: float(r[h.index(name)])
    cyc = g("sm__cycles_elapsed.max") if "sm__cycles_elapsed.max" in h else g("smsp__cycles_elapsed.max")
    dadd = g("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed") * cyc
    dmul = g("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed") * cyc
    dfma = g("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed") * cyc
    unit = lambda name: {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}
    def bytes_(name):
        i = h.index(name)
        u = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout.splitlines()))[1][i]
        return float(r[i]) * {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}.get(u, 1.0)
    return {"dadd": dadd, "dmul": dmul, "dfma": dfma, "flops": dadd + dmul + 2 * dfma, "cycles": cyc,
            "dram_bytes": bytes_("dram__bytes_read.sum") + bytes_("dram__bytes_write.sum"),
            "duration_us": g("gpu__time_duration.sum") * (1e3 if g("gpu__time_duration.sum") < 20 else 1.0),
            "inst": g("smsp__inst_executed.sum"), "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "fp64_pipe_pct": g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), "regs": g("launch__registers_per_thread"),
            "grid": g("launch__grid_size")}


one, sat = counters(sys.argv[1], "lm2_kernel"), counters(sys.argv[2], "lm2_kernel")
ev1, evs = float(sys.argv[3]), float(sys.argv[4])
bm = counters(sys.argv[1], "bm_.*kernel")
out = {
    "source": "ncu --set full --clock-control none on `bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --no-extras` (one frame's launch, "
              "gpurun_out/prof_r2.ncu-rep) and on `scripts/lm_saturation.py --child 12011` (16 frames' seeds in one launch, gpurun_out/lm_sat_r2.ncu-rep); "
              "thread-level smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on x elapsed cycles; flops = dadd + dmul + 2 dfma",
    "kernel": "lm2_kernel<7,20>",
    "fp64_flops_per_launch": one["flops"], "evals_executed_per_launch": ev1, "fp64_flops_per_executed_eval": one["flops"] / ev1,
    "dram_bytes_per_launch": one["dram_bytes"], "inst_per_launch": one["inst"], "registers": one["regs"],
    "one_frame": {"duration_us": one["duration_us"], "issue_active_pct": one["issue_active_pct"], "fp64_pipe_pct": one["fp64_pipe_pct"], "grid": one["grid"]},
    "saturated": {"duration_us": sat["duration_us"], "evals": evs, "fp64_flops": sat["flops"], "fp64_tflops": sat["flops"] / (sat["duration_us"] * 1e-6) / 1e12,
                  "issue_active_pct": sat["issue_active_pct"], "fp64_pipe_pct": sat["fp64_pipe_pct"], "grid": sat["grid"],
                  "frames_per_launch": 16, "ms_per_frame": sat["duration_us"] / 16 / 1e3,
                  "note": "16 frames' seeds in one launch: no tail, the kernel's sustained rate"},
    "bm_dram_bytes_per_launch": bm["dram_bytes"],
}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_lm_counters.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
