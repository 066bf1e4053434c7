"""The bench.py JSON contract, checked on the lines recorded from the last B200 run (profiles/r1_bench.jsonl) and on the
parts of bench.py that run without a GPU (argument parsing, workload generator, CPU reference leg on one frame)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lines():
    with open(os.path.join(ROOT, "profiles", "r1_bench.jsonl")) as f:
        return [json.loads(l) for l in f if l.strip()]


def test_recorded_lines_follow_the_contract():
    lines = _lines()
    ours = [d for d in lines if d.get("impl", "ours") != "reference"]
    ref = [d for d in lines if d.get("impl") == "reference"]
    assert ours and ref
    for d in ours:
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                  "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
            assert k in d, k
        assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
        assert "workload" in d["config"] and "model" not in d["config"]
        assert d["warmup"] >= 3 and d["gpu_launches"] > 0
        e = d["e2e"]
        assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
        c = d["clocks"]
        assert c["sm_mhz"] and c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        r = d["roofline"]
        assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["unit"] == "GB/s"
        executed = d["per_step"]["bm_evals"] + d["per_step"]["lm_evals_executed"]
        # value = executed evaluations / time (per-GPU work differs only for distinct streams)
        assert d["n_gpus"] > 1 or abs(d["value"] * d["ms_per_step"] * 1e-3 - executed) < 1.0
    one = [d for d in ours if d["n_gpus"] == 1][0]
    cb = one["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0 and "sample" in cb
    assert one["value"] > 1e8                                  # BASELINE.json target of the north star
    for d in ref:
        assert d["impl"] == "reference" and d["metric"] == one["metric"] and d["unit"] == one["unit"]
        assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
        assert d["cpu_baseline"]["value"] == d["value"]


def _r2_lines():
    path = os.path.join(ROOT, "profiles", "r2_bench.jsonl")
    if not os.path.exists(path):
        return []
    with open(path) as f:
        return [json.loads(l) for l in f if l.strip()]


def test_round2_lines_follow_the_contract():
    """Round-2 format: the roofline of the dominant kernel is measured against the FP64 rate (what ncu shows as the bound), on an
    isolated launch, with the peak measured in the same run; the HBM view sits under roofline.hbm; the line carries the parity
    block of the benchmarked configuration and the shortcut-matched CPU arm."""
    lines = _r2_lines()
    if not lines:
        import pytest
        pytest.skip("no round-2 bench lines recorded yet")
    ours = [d for d in lines if d.get("impl", "ours") != "reference"]
    ref = [d for d in lines if d.get("impl") == "reference"]
    assert ours and ref
    for d in ours:
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                  "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "timing"):
            assert k in d, k
        assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
        assert "workload" in d["config"] and "model" not in d["config"]
        assert d["warmup"] >= 3 and d["gpu_launches"] > 0
        e = d["e2e"]
        assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
        c = d["clocks"]
        assert c["sm_mhz"] and c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        executed = d["per_step"]["bm_evals"] + d["per_step"]["lm_evals_executed"]
        assert abs(d["value"] * d["ms_per_step"] * 1e-3 - executed * d["n_gpus"]) < 1.0 * d["n_gpus"]
    one = [d for d in ours if d["n_gpus"] == 1 and "roofline" in d and "cpu_baseline" in d][0]
    assert one["timing"]["timed_s_total"] >= 0.5                     # timed region >= 0.5 s
    r = one["roofline"]
    assert r["bound"] == "fp64" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["ms_per_launch"] > one["ms_per_step"] * 0.5             # an isolated launch, not a pipelined stage latency
    assert 20.0 < r["peak"] < 45.0 and r["traffic"] and r["traffic"] < r["hbm"]["algorithmic_bytes_per_launch"]
    h = r["hbm"]
    assert h["bound"] == "hbm" and h["unit"] == "GB/s" and abs(h["frac"] - h["achieved"] / h["peak"]) < 1e-12
    o = one["roofline_other"]
    assert o["bound"] == "hbm" and abs(o["frac"] - o["achieved"] / o["peak"]) < 1e-12
    cb = one["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0 and "sample" in cb
    assert cb["with_shortcut"]["value"] > cb["value"]                # the shortcut-matched arm is the faster CPU baseline
    assert one["value"] > 1e8
    par = one["parity"]
    assert par["frames_checked"] >= 3 and par["disparity_mismatches"] == 0 and par["accept_set_symmetric_difference"] == 0
    assert par["map_order_equal"] and par["inv_depth_l1"] < 1e-6
    for d in ref:
        assert d["impl"] == "reference" and d["metric"] == one["metric"] and d["unit"] == one["unit"]
        assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
        assert d["cpu_baseline"]["value"] == d["value"]


def test_bench_cli_and_cpu_leg():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True)
    assert p.returncode == 0 and "--impl" in p.stdout and "--pipeline-depth" in p.stdout
    sys.path.insert(0, ROOT)
    import bench
    base = bench.make_workload(seed=10)
    f1 = bench.shifted(base, 3)
    assert f1["t_ts_ns"] - base["t_ts_ns"] == 3 * int(round(bench.FRAME_MS * 1e6))
    assert np.array_equal(f1["left"]["x"], base["left"]["x"]) and (f1["left"]["t"] > base["left"]["t"]).all()
    leg = bench.cpu_leg(base, sample_steps=1)
    assert leg["kind"] == "port" and leg["value"] > 1e4 and leg["cores"] >= 1
