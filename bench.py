#!/usr/bin/env python
"""bench.py -- headline benchmark of the ESVO mapping hot path on B200.

Metric (BASELINE.json): depth-candidate patch evaluations / s (EventBM zncc_cost evaluations + DepthProblem::operator()
evaluations actually executed) on BASELINE.json configs[1]: 346x260 stereo time-surface pair + 5k DepthProblem seeds,
one event stream per GPU (weak scaling).

One "step" = one mapping frame of one stream, everything the reference does between two MappingAtTime calls: ingest the
frame's raw events of both cameras, build both time surfaces at the frame stamp, hand them to the mapper, block-match the
5000 newest left events, LM-refine the matches, cull, update the fusion window, fuse the whole window, clean, regularise.

  value   : device-timed (CUDA events) with every input already resident in HBM (esvo_*_dev entry points), max over ranks;
  e2e     : the same step through the host-buffer C ABI from pinned host memory, H2D of the events / seeds / poses and D2H of
            the counters and of the ordered fused map inside the timed region;
  K steps are timed as ONE region (barrier + synchronise on both sides); the region is repeated R times on fresh stamps
  until >= 0.5 s have been timed and the MEDIAN region is reported (`steps` stays K; `timing` says how many regions);
  parity  : (1 GPU) the same frames through the CPU oracle with the product's own rectification tables, outside the timed
            regions: idx-grid / TS equality, accept-set difference, disparity equality, inverse-depth L1, fused-map order;
  extras  : (1 GPU) BASELINE configs[2] (640x480, 20k seeds + fusion) and configs[3] (tracking ms/frame, pose vs oracle);
  --impl reference : the CPU restatement of the reference (oracle/, all host threads) on the same workload -- the only
            legs where oracle/ is executed are this one, cpu_baseline and parity.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the library uses one stream per pipeline slot + 3 service streams per GPU: give every stream its own hardware queue
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

BM_BYTES = 420.0          # 105 px x 4 B per BM candidate (f32 TS convention)
LM_BYTES = 1024.0         # 2 x (15+1)(7+1) px x 4 B per LM residual evaluation
FRAME_MS = 50.0           # mapping_rate_hz = 20
N_SEEDS = 5000
RIG = "hkust"
CONFIGS = {
    "cfg2": dict(rig="hkust", n_seeds=5000, synth={},
                 name="BASELINE configs[1]: 346x260 stereo TS pair + 5k DepthProblem seeds per frame (hkust rig, mapping_hkust.yaml)"),
    "cfg3": dict(rig="dsec", n_seeds=20000, synth=dict(n_segments=120),
                 name="BASELINE configs[2]: 640x480 (DSEC-shaped) stereo TS, 20k seeds + DepthFusion (dsec rig, mapping_dsec.yaml: "
                      "fusion_radius 1, 5-frame window, SmoothTimeSurface)"),
}
METRIC = "depth-candidate patch evals/s (EventBM zncc + DepthProblem LM) @346x260, 5k events/frame into BM"
CTR_KEYS = ["n_events", "n_seeds", "n_solved", "n_culled", "n_fusions", "bm_evals", "lm_evals", "map_size"]


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def load_ncu_counters():
    """Per-launch counters of the dominant kernel from the committed ncu capture of this workload (profiles/): FP64 operations
    per executed evaluation and DRAM bytes per launch.  Absent file -> None (the line then carries null, never a literal)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_lm_counters.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed regions (in-process NVML; an `nvidia-smi -lms` loop
    re-enumerates the device under the driver lock and measurably stalled kernel submission)."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self._stop_evt, self.proc = gpu, [], threading.Event(), None
        self.period = float(os.environ.get("ESVO_BENCH_SMI_MS", "25")) / 1e3
        self.windows, self.stamps = [], []
        self.mode = os.environ.get("ESVO_BENCH_SAMPLER", "nvml")

    def _run_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        idx = self.gpu
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        if vis and all(v.strip().isdigit() for v in vis.split(",")) and self.gpu < len(vis.split(",")):
            idx = int(vis.split(",")[self.gpu])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        while not self._stop_evt.is_set():
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            try:
                mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.rows.append([str(sm), str(mx)] + ["Active" if mask & bit else "Not Active" for bit, _ in self.REASONS])
            self.stamps.append(time.perf_counter())
            self._stop_evt.wait(self.period)

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-lms", str(max(1, int(self.period * 1e3)))], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])
            self.stamps.append(time.perf_counter())
            if self._stop_evt.is_set():
                break

    def run(self):
        try:
            if self.mode == "nvml":
                try:
                    self._run_nvml()
                    return
                except Exception:
                    if self.rows:
                        return
                    self.mode = "smi"
            self._run_smi()
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        if self.proc:
            self.proc.terminate()
        rows = self.rows
        if self.windows:
            inside = [r for r, t in zip(self.rows, self.stamps) if any(a <= t <= b for a, b in self.windows)]
            rows = inside or rows
        sm = [int(r[0]) for r in rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        for r in rows:
            for (_, name), v in zip(self.REASONS, r[2:6]):
                if v == "Active":
                    reasons.add(name)
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "sampler": self.mode,
                "scope": "samples inside the timed regions" if rows is not self.rows else "whole run"}


def make_workload(seed, cfg="cfg2"):
    from esvo_b200 import synth
    c = CONFIGS[cfg]
    # one 50 ms mapping frame of the stream; the frame is replayed with shifted stamps every step
    return synth.make_stream(c["rig"], seed=seed, n_seeds=c["n_seeds"], history_ms=FRAME_MS, **c["synth"])


def shifted(s, k):
    """Frame k of the synthetic stream = the base frame with all stamps advanced by k frame periods."""
    dt = int(round(FRAME_MS * 1e6)) * k
    out = {"t_ts_ns": s["t_ts_ns"] + dt, "T_world_left": s["T_world_left"], "poses": s["poses"],
           "pose_t": s["pose_t"] + dt}
    for side in ("left", "right"):
        e = s[side]
        out[side] = dict(x=e["x"], y=e["y"], t=e["t"] + dt, p=e["p"])
    sd = s["seeds"]
    out["seeds"] = dict(x=sd["x"], y=sd["y"], t=sd["t"] + dt)
    return out


u16, i64, u8, f64 = C.POINTER(C.c_uint16), C.POINTER(C.c_int64), C.POINTER(C.c_uint8), C.POINTER(C.c_double)
FRAME_NS = int(round(FRAME_MS * 1e6))


class StreamDriver:
    """Feeds frames of one synthetic stream to one esvo ctx, either from HBM-resident arrays or from pinned host arrays."""

    def __init__(self, g, depth):
        self.g, self.depth = g, depth
        self.tickets = []
        self.h2d = self.d2h = 0
        self.last_res = None

    @staticmethod
    def to_dev(f):
        import torch
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        d = {side: {k: dev(v) for k, v in f[side].items()} for side in ("left", "right")}
        d["seeds"] = {k: dev(v) for k, v in f["seeds"].items()}
        d["pose_t"] = dev(f["pose_t"]); d["poses"] = dev(f["poses"])
        d["t_ts_ns"] = int(f["t_ts_ns"]); d["T"] = np.ascontiguousarray(f["T_world_left"], np.float64)
        return d

    @staticmethod
    def to_pinned(f):
        import torch
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

        def packet(e):
            """One pinned packet buffer per camera and frame, laid out t | x | y | p like a driver's SoA event packet: the library
            moves arrays that are adjacent in host memory with a single copy (esvo_stage_ts_events)."""
            n = e["x"].size
            buf = torch.zeros(13 * n, dtype=torch.uint8).pin_memory().numpy()
            out = {"t": buf[:8 * n].view(np.int64), "x": buf[8 * n:10 * n].view(np.uint16), "y": buf[10 * n:12 * n].view(np.uint16), "p": buf[12 * n:13 * n]}
            for k in out:
                out[k][:] = e[k]
            return out
        d = {side: packet(f[side]) for side in ("left", "right")}
        d["seeds"] = {k: pin(v) for k, v in f["seeds"].items()}
        d["pose_t"] = pin(f["pose_t"]); d["poses"] = pin(f["poses"])
        d["t_ts_ns"] = int(f["t_ts_ns"]); d["T"] = np.ascontiguousarray(f["T_world_left"], np.float64)
        return d

    @staticmethod
    def advance(d, nframes):
        """Move a frame's stamps forward by nframes frame periods, in place (device tensors or pinned numpy arrays)."""
        dt = FRAME_NS * nframes
        for side in ("left", "right"):
            d[side]["t"] += dt
        d["seeds"]["t"] += dt
        d["pose_t"] += dt
        d["t_ts_ns"] += dt

    def step_resident(self, d):
        """One step, every input already resident in HBM."""
        g = self.g
        P = lambda t, ty: C.cast(t.data_ptr(), ty)
        for cam, side in ((0, "left"), (1, "right")):
            e = d[side]
            g._call("ts_push_events_dev", [C.c_int, u16, u16, i64, u8, C.c_size_t], cam, P(e["x"], u16), P(e["y"], u16),
                    P(e["t"], i64), P(e["p"], u8), e["x"].numel())
            g.run_ts_build(cam, d["t_ts_ns"])
        g._call("set_ts_pair_dev", [f64], d["T"].ctypes.data_as(f64))
        sd = d["seeds"]
        g._call("stage_mapping_inputs_dev", [u16, u16, i64, C.c_size_t, i64, f64, C.c_size_t], P(sd["x"], u16), P(sd["y"], u16),
                P(sd["t"], i64), sd["x"].numel(), P(d["pose_t"], i64), P(d["poses"], f64), d["pose_t"].numel())
        g.run_mapping()

    def step_e2e(self, d):
        """One step through the host-buffer C ABI: pinned host arrays in, ordered fused map + counters out (collected
        depth-1 frames later so that frames stay in flight)."""
        g = self.g
        if len(self.tickets) >= max(1, self.depth - 1):
            self._collect()
        for cam, side in ((0, "left"), (1, "right")):
            e = d[side]
            g.stage_ts_events(cam, e["x"], e["y"], e["t"], e["p"])
            g.run_ts_build(cam, d["t_ts_ns"])
            self.h2d += e["x"].size * 13
        g._call("set_ts_pair_dev", [f64], d["T"].ctypes.data_as(f64))
        sd = d["seeds"]
        g.stage_mapping_inputs(sd["x"], sd["y"], sd["t"], d["pose_t"], d["poses"])
        g.run_mapping()
        self.tickets.append(g.results_begin())
        self.h2d += sd["x"].size * 12 + d["pose_t"].size * 136 + 128

    def _collect(self):
        m, res = self.g.results_end_view(self.tickets.pop(0))
        self.d2h += 64 + 128 + m.nbytes
        self.last_res = res

    def drain_e2e(self):
        while self.tickets:
            self._collect()


def measure_stream(args, g, base, world, local_rank, sampler, K, Wm, cfg_params, want_breakdown=True, target_s=0.5, max_regions=400):
    """Priming + warm-up + R timed regions of K steps for both legs on ctx g.  Returns a dict of raw measurements."""
    import torch
    from esvo_b200 import dist as edist
    NP = cfg_params.max_num_fusion_frames
    drv = StreamDriver(g, args.pipeline_depth)
    stream = torch.cuda.ExternalStream(g.stream(), device=local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    KP = min(K, 24) if want_breakdown else 0
    k_next = 0
    prime = [drv.to_dev(shifted(base, k)) for k in range(NP)]; k_next += NP
    prof = [drv.to_dev(shifted(base, k_next + k)) for k in range(KP)]; k_next += KP
    warm = [drv.to_dev(shifted(base, k_next + k)) for k in range(Wm)]; k_next += Wm
    steps = [drv.to_dev(shifted(base, k_next + k)) for k in range(K)]
    torch.cuda.synchronize()
    ms = (C.c_double * 8)(); cnt = (C.c_uint64 * 8)()
    out = {}
    with torch.cuda.stream(stream):
        for d in prime:                       # fill the fusion window (a tracker that has been running for a while)
            drv.step_resident(d)
        g.sync()
        if KP:                                # per-stage breakdown: a separately profiled pass, NOT part of `value`
            g._call("profile", [C.c_int], 0xFF)
            g._call("profile_read", [f64, C.POINTER(C.c_uint64)], (C.c_double * 8)(), (C.c_uint64 * 8)())
            for d in prof:
                drv.step_resident(d)
            g.sync()
            if os.environ.get("ESVO_BENCH_TIMELINE"):
                g._call("profile_dump", [C.c_char_p], os.environ["ESVO_BENCH_TIMELINE"].encode())
            g._call("profile_read", [f64, C.POINTER(C.c_uint64)], ms, cnt)
            g._call("profile", [C.c_int], 0)
        for d in warm:
            drv.step_resident(d)
        g.sync()
        flush.fill_(7)            # evict everything (inputs of the timed steps included) from L2 before the first timed region
        barrier()
        # ---------------- leg 1: inputs resident in HBM, R regions of K steps ----------------
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        region_ms, issue = [], []
        launches0 = g.launch_count()
        R = None
        r = 0
        while True:
            if r > 0:
                for d in steps:
                    drv.advance(d, K)
                torch.cuda.synchronize()
            barrier()
            t0w = time.perf_counter()
            e0.record(stream)
            for d in steps:
                drv.step_resident(d)
            t_issue = time.perf_counter() - t0w
            g.sync()                                       # every stream of the library drained
            e1.record(stream)
            barrier()
            sampler.windows.append((t0w, time.perf_counter()))
            region_ms.append(e0.elapsed_time(e1)); issue.append(t_issue * 1e3 / K)
            if r == 0:
                out["launches_per_region"] = g.launch_count() - launches0
            r += 1
            if R is None:      # number of regions: agreed across ranks from the first one
                want = int(np.clip(np.ceil(1.15 * target_s * 1e3 / max(region_ms[0], 1e-3)) + 1, 3, max_regions))   # the first region runs slower than the rest
                R = int(edist.gather_scalars([float(want)], device="cuda")[:, 0].max())
            if r >= R:
                break
        ctr = g.fetch_mapping_counters()
        dbg = g.L.lib.esvo_debug_counter
        dbg.argtypes = [C.c_void_p, C.c_int]; dbg.restype = C.c_uint64
        out["lm_exec"] = int(dbg(g.ctx, 7))
        out["ctr"] = ctr
    out["region_ms"] = region_ms; out["issue_resident_ms"] = issue
    out["breakdown"] = {"time_surface_x2": ms[0] / KP, "block_matching": ms[1] / KP, "seed_order": ms[2] / KP, "depth_lm": ms[3] / KP,
                        "point_order_cull": ms[4] / KP, "fusion_clean_regularise": ms[5] / KP,
                        "source": f"separate profiled pass of {KP} steps before the warm-up (stage durations with other pipeline slots "
                                  "running: they overlap, the sum exceeds ms_per_step)"} if KP else None
    # ---------------- TS-only rate (second half of the metric): resident events -> rectified u8 TS, both cameras ----------------
    with torch.cuda.stream(stream):
        for d in steps:
            drv.advance(d, K)
        torch.cuda.synchronize()
        e0.record(stream)
        for d in steps:
            for cam, side in ((0, "left"), (1, "right")):
                e = d[side]
                g._call("ts_push_events_dev", [C.c_int, u16, u16, i64, u8, C.c_size_t], cam, C.cast(e["x"].data_ptr(), u16),
                        C.cast(e["y"].data_ptr(), u16), C.cast(e["t"].data_ptr(), i64), C.cast(e["p"].data_ptr(), u8), e["x"].numel())
                g.run_ts_build(cam, d["t_ts_ns"])
        g.sync()
        e1.record(stream)
        torch.cuda.synchronize()
        out["ts_only_frames_per_s"] = 2 * K / (e0.elapsed_time(e1) * 1e-3)
    # ---------------- leg 2: end to end through the host-buffer C ABI ----------------
    R2 = len(region_ms)
    consumed = k_next + K * (R2 + 1)          # frames consumed so far on the stream's time line (regions + the TS-only pass)
    pw = [drv.to_pinned(shifted(base, consumed + k)) for k in range(Wm)]
    ps = [drv.to_pinned(shifted(base, consumed + Wm + k)) for k in range(K)]
    for d in pw:
        drv.step_e2e(d)
    drv.drain_e2e()
    e2e_ms, issue_e = [], []
    for r in range(R2):
        if r > 0:
            for d in ps:
                drv.advance(d, K)
        drv.h2d = drv.d2h = 0
        barrier()
        t0 = time.perf_counter()
        for d in ps:
            drv.step_e2e(d)
        t_issue = time.perf_counter() - t0
        drv.drain_e2e()
        barrier()
        t1 = time.perf_counter()
        sampler.windows.append((t0, t1))
        e2e_ms.append((t1 - t0) * 1e3); issue_e.append(t_issue * 1e3 / K)
    out["e2e_region_ms"] = e2e_ms; out["issue_e2e_ms"] = issue_e
    out["h2d_per_step"] = drv.h2d // K; out["d2h_per_step"] = drv.d2h // K
    out["e2e_ctr"] = drv.last_res
    return out


def isolated_kernels(g, base):
    """One un-overlapped launch of BM and of the depth LM on the frame's own data (pipeline idle), CUDA-event timed inside the
    library (esvo_profile stages 1 / 3): the kernel-efficiency numbers of the roofline."""
    g.sync()
    g._call("set_pipeline_depth", [C.c_int], 1)
    for cam in (0, 1):
        g.ts_reset(cam)
    f = shifted(base, 0)
    for cam, side in ((0, "left"), (1, "right")):
        e = f[side]
        g.ts_push_events(cam, e["x"], e["y"], e["t"], e["p"])
        g.ts_build(cam, f["t_ts_ns"], want_idx=False, want_ts=False)
    g.set_ts_pair(None, None, f["T_world_left"])
    g._call("profile", [C.c_int], (1 << 1) | (1 << 3))

    def rd():
        m, c = (C.c_double * 8)(), (C.c_uint64 * 8)()
        g._call("profile_read", [f64, C.POINTER(C.c_uint64)], m, c)
        return list(m)
    rd()
    sd = f["seeds"]
    bm_t, lm_t = [], []
    for _ in range(4):
        seeds, bm_evals = g.bm_match(sd["x"], sd["y"], sd["t"], f["pose_t"], f["poses"])
        bm_t.append(rd()[1])
        pts, nfev = g.depth_solve(seeds)
        lm_t.append(rd()[3])
    g._call("profile", [C.c_int], 0)
    dbg = g.L.lib.esvo_debug_counter
    dbg.argtypes = [C.c_void_p, C.c_int]; dbg.restype = C.c_uint64
    return {"bm_ms": float(min(bm_t[1:])), "lm_ms": float(min(lm_t[1:])), "n_seeds": int(seeds.size), "bm_evals": int(bm_evals),
            "lm_nfev": int(nfev), "lm_exec": int(dbg(g.ctx, 7)), "n_events": int(sd["x"].size)}


def parity_block(prod, base, rig, n_check=3, tracking=True):
    """The benchmarked frames through the product (fresh ctx, depth 1) and through the CPU oracle, both on the PRODUCT's own
    rectification tables: fusion window primed to its full length, then n_check consecutive frames compared stage by stage."""
    from esvo_b200 import capi, configs, dist as edist
    from oracle.loader import load_oracle
    orc = load_oracle()
    l, r = configs.rig_calibs(rig)
    pp, po = configs.params_for(rig, prod), configs.params_for(rig, orc)
    g = capi.Backend(prod, l, r, pp)
    o = capi.Backend(orc, l, r, po)
    orc.lib.esvo_oracle_set_exec_threads(o.ctx, C.c_int(os.cpu_count() or 1))
    for cam in (0, 1):
        o.set_rectify_tables(cam, *g.get_rectify_tables(cam))
    NP = pp.max_num_fusion_frames
    res = {"frames_primed": NP, "frames_checked": n_check, "tables": "product's own (esvo_compute_rectify_tables, pinned to cv2)",
           "idx_grid_equal": True, "ts_bytes_equal": True, "ts_mismatching_pixels": 0, "accept_set_symmetric_difference": 0,
           "disparity_mismatches": 0, "seeds_compared": 0, "lm_points_compared": 0, "lm_accept_symmetric_difference": 0,
           "inv_depth_l1": 0.0, "inv_depth_max_rel": 0.0, "map_order_equal": True, "map_size_equal": True, "map_inv_depth_l1": 0.0,
           "map_inv_depth_max_rel": 0.0, "map_checksum_equal": True, "counters_equal": True}
    l1, l1m = [], []
    last = None
    t0 = time.perf_counter()
    for k in range(NP + n_check):
        f = shifted(base, k)
        check = k >= NP
        for be in (g, o):
            for cam, side in ((0, "left"), (1, "right")):
                e = f[side]
                be.ts_push_events(cam, e["x"], e["y"], e["t"], e["p"])
        if check:
            for cam in (0, 1):
                ig, tg = g.ts_build(cam, f["t_ts_ns"]); io, to = o.ts_build(cam, f["t_ts_ns"])
                res["idx_grid_equal"] &= bool(np.array_equal(ig, io))
                res["ts_bytes_equal"] &= bool(np.array_equal(tg, to))
                res["ts_mismatching_pixels"] += int((tg != to).sum())
        else:
            for be in (g, o):
                for cam in (0, 1):
                    be.ts_build(cam, f["t_ts_ns"], want_idx=False, want_ts=False)
        for be in (g, o):
            be.set_ts_pair(None, None, f["T_world_left"])
        sd = f["seeds"]
        if check:
            sg, evg = g.bm_match(sd["x"], sd["y"], sd["t"], f["pose_t"], f["poses"])
            so, evo = o.bm_match(sd["x"], sd["y"], sd["t"], f["pose_t"], f["poses"])
            kg = {(float(a["x_left_raw"][0]), float(a["x_left_raw"][1]), int(a["t_ns"])): float(a["disp"]) for a in sg}
            ko = {(float(a["x_left_raw"][0]), float(a["x_left_raw"][1]), int(a["t_ns"])): float(a["disp"]) for a in so}
            res["accept_set_symmetric_difference"] += len(set(kg) ^ set(ko))
            common = set(kg) & set(ko)
            res["disparity_mismatches"] += sum(1 for q in common if kg[q] != ko[q])
            res["seeds_compared"] += len(common)
            res["counters_equal"] &= (evg == evo)
            pg, _ = g.depth_solve(so)          # the SAME seeds into both solvers
            po_, _ = o.depth_solve(so)
            key = lambda p: (float(p["x"][0]), float(p["x"][1]), float(p["T_world_cam"][3]), float(p["T_world_cam"][7]), float(p["T_world_cam"][11]))
            dg = {key(p): p for p in pg}; do = {key(p): p for p in po_}
            res["lm_accept_symmetric_difference"] += len(set(dg) ^ set(do))
            cm = [q for q in do if q in dg]
            rel = np.array([abs(dg[q]["inv_depth"] - do[q]["inv_depth"]) / abs(do[q]["inv_depth"]) for q in cm]) if cm else np.zeros(0)
            l1.append(rel); res["lm_points_compared"] += len(cm)
        cg = g.mapping_at_time(sd["x"], sd["y"], sd["t"], f["pose_t"], f["poses"])
        co = o.mapping_at_time(sd["x"], sd["y"], sd["t"], f["pose_t"], f["poses"])
        if check:
            mg, mo = g.map_download(), o.map_download()
            res["counters_equal"] &= all(cg[q] == co[q] for q in ("n_events", "n_seeds", "n_solved", "n_culled", "n_fusions", "bm_evals", "map_size"))
            res["map_size_equal"] &= (mg.size == mo.size)
            same_order = mg.size == mo.size and np.array_equal(mg["row"], mo["row"]) and np.array_equal(mg["col"], mo["col"])
            res["map_order_equal"] &= bool(same_order)
            if same_order:
                v = mo["inv_depth"] > -1e-6
                relm = np.abs(mg["inv_depth"][v] - mo["inv_depth"][v]) / np.abs(mo["inv_depth"][v])
                l1m.append(relm)
                res["map_order_equal"] &= bool(np.array_equal(mg["inv_depth"] > -1e-6, v) and np.array_equal(mg["age"], mo["age"]))
            ck_g, ck_o = edist.map_checksum(mg), edist.map_checksum(mo)
            res["map_checksum_equal"] &= bool(abs(ck_g - ck_o) <= 1e-9 * max(1.0, abs(ck_o)))
            last = (mg, mo, f)
    if l1:
        a = np.concatenate(l1)
        res["inv_depth_l1"] = float(a.mean()) if a.size else 0.0; res["inv_depth_max_rel"] = float(a.max()) if a.size else 0.0
        res["inv_depth_frac_above_1e-4"] = float((a > 1e-4).mean()) if a.size else 0.0
    if l1m:
        a = np.concatenate(l1m)
        res["map_inv_depth_l1"] = float(a.mean()) if a.size else 0.0; res["map_inv_depth_max_rel"] = float(a.max()) if a.size else 0.0
        res["map_inv_depth_frac_above_1e-4"] = float((a > 1e-4).mean()) if a.size else 0.0
    res["seconds"] = time.perf_counter() - t0
    extra = tracking_block(g, o, pp, last) if (tracking and last is not None) else None
    g.close(); o.close()
    return res, extra


def tracking_block(g, o, prm, last):
    """BASELINE configs[3] (tracking half of the loop): the fused local map of the last checked frame is handed to the tracker the
    way the nodes do it (world-frame f32 point cloud), and the next time surface (10 ms later: tracking runs at 100 Hz) is tracked
    with RegProblemLM (tracking_hkust.yaml: 500-point batches, <= 2000 points, Huber 50).  ms/frame = reset + solve, host wall
    clock around the two synchronous C-ABI calls; pose vs the oracle's on the same inputs."""
    from esvo_b200 import synth
    mg, mo, f = last
    Tw = np.asarray(f["T_world_left"], float)
    cloud = (mg["p_cam"] @ Tw[:3, :3].T + Tw[:3, 3]).astype(np.float32)
    if cloud.shape[0] < prm.trk_batch_size:
        return {"skipped": f"local map has {cloud.shape[0]} points < batch size"}
    out = {"ref_points": int(cloud.shape[0]), "batch": int(prm.trk_batch_size)}
    s2 = synth.make_stream(RIG, seed=10, n_seeds=100, history_ms=FRAME_MS, t_ts=0.51)
    for be in (g, o):
        be.ts_reset(0)
        e = s2["left"]
        be.ts_push_events(0, e["x"], e["y"], e["t"], e["p"])
    _, ts_cur = g.ts_build(0, s2["t_ts_ns"], want_idx=False)
    _, ts_cur_o = o.ts_build(0, s2["t_ts_ns"], want_idx=False)
    out["ts_equal"] = bool(np.array_equal(ts_cur, ts_cur_o))
    T_prior = np.asarray(s2["T_world_left"], float)
    for name, analytical in (("analytical", True), ("numerical", False)):
        poses = {}
        for tag, be in (("gpu", g), ("cpu", o)):
            c = cloud.copy()
            be.track_srand(1)
            be.track_reset(c, Tw, Tw, ts_cur)
            poses[tag] = be.track_solve(analytical)
        Tg, To = poses["gpu"][0], poses["cpu"][0]
        tms, cms = [], []
        for rep in range(12):
            c = cloud.copy(); g.track_srand(1)
            t0 = time.perf_counter(); g.track_reset(c, Tw, Tw, ts_cur); t1 = time.perf_counter(); g.track_solve(analytical); t2 = time.perf_counter()
            tms.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3))
        for rep in range(2):
            c = cloud.copy(); o.track_srand(1)
            t0 = time.perf_counter(); o.track_reset(c, Tw, Tw, ts_cur); o.track_solve(analytical); cms.append((time.perf_counter() - t0) * 1e3)
        tms = np.array(tms[2:])
        dR = Tg[:3, :3] @ To[:3, :3].T
        ang = float(np.degrees(np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1))))
        gt = float(np.linalg.norm(Tg[:3, 3] - T_prior[:3, 3]))
        out[name] = {"ms_per_frame": float(np.median(tms.sum(axis=1))), "reset_ms": float(np.median(tms[:, 0])), "solve_ms": float(np.median(tms[:, 1])),
                     "cpu_oracle_ms_per_frame": float(min(cms)), "stats_equal": poses["gpu"][1] == poses["cpu"][1], "stats": poses["gpu"][1],
                     "pose_max_abs_diff_vs_oracle": float(np.abs(Tg - To).max()), "rotation_diff_deg_vs_oracle": ang,
                     "translation_rel_diff_vs_oracle": float(np.linalg.norm(Tg[:3, 3] - To[:3, 3]) / max(np.linalg.norm(To[:3, 3]), 1e-12)),
                     "translation_error_vs_ground_truth_m": gt}
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    from esvo_b200 import capi, configs
    from esvo_b200 import dist as edist
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        # NCCL may print its version banner on stdout; the contract is ONE JSON line there
        os.environ["NCCL_DEBUG"] = os.environ.get("NCCL_DEBUG", "WARN") if os.environ.get("NCCL_DEBUG", "").upper() not in ("VERSION", "") else "WARN"
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    prod = capi.load_product()
    cfg = CONFIGS["cfg2"]
    l, r = configs.rig_calibs(cfg["rig"])
    prm = configs.params_for(cfg["rig"], prod)
    g = capi.Backend(prod, l, r, prm, device=local_rank)
    g._call("set_pipeline_depth", [C.c_int], args.pipeline_depth)
    sampler = ClockSampler(local_rank); sampler.start()
    t_wait = time.time()
    while not sampler.rows and time.time() - t_wait < 15:
        time.sleep(0.05)
    # Weak scaling needs the same work on every GPU: the headline replays the SAME synthetic stream shape on every rank (scene
    # seed 10; time origin differs per rank).  `--streams distinct` gives every rank its own scene (seeds 10, 11, ... as in
    # SURVEY cfg 5); under N > 1 a short distinct-scenes pass is ALSO run and reported as extras.distinct_streams.
    base = make_workload(seed=edist.stream_seed(rank if args.streams == "distinct" else 0, base=args.scene_seed))
    if args.streams != "distinct" and rank:
        base = shifted(base, 100000 * rank)
    K, Wm = args.steps, args.warmup
    m = measure_stream(args, g, base, world, local_rank, sampler, K, Wm, prm, target_s=args.min_timed_s)
    ctr, lm_exec = m["ctr"], m["lm_exec"]
    evals_step = ctr["bm_evals"] + lm_exec
    ce = m["e2e_ctr"]
    # Both legs count the evaluations actually EXECUTED (the LM kernel re-uses f(x) inside the forward difference where
    # NumericalDiff recomputes it).  The result hand-off of the e2e leg carries the nfev-style counter only; the frames of both
    # legs are identical in content, so the executed count per frame is the resident leg's (checked through the nfev counter).
    e2e_evals = ce["bm_evals"] + (lm_exec if ce["lm_evals"] == ctr["lm_evals"] else ce["lm_evals"])
    reg = edist.gather_scalars(m["region_ms"], device="cuda").max(axis=0)          # per region: max over ranks
    reg_e = edist.gather_scalars(m["e2e_region_ms"], device="cuda").max(axis=0)
    med_ms, med_e = float(np.median(reg)), float(np.median(reg_e))
    m_last = g.map_download()
    rec = edist.make_record(edist.stream_seed(rank), K, ce, edist.map_checksum(m_last))
    _, evals_all, records = edist.reduce_and_gather(med_ms, float(evals_step), rec, device="cuda")
    _, e2e_evals_all, _ = edist.reduce_and_gather(med_e, float(e2e_evals), rec, device="cuda")
    per_rank = edist.gather_scalars([float(np.median(m["region_ms"])) / K, float(np.median(m["e2e_region_ms"])) / K,
                                     float(np.mean(m["issue_resident_ms"])), float(np.mean(m["issue_e2e_ms"]))], device="cuda")
    # isolated launches + FP64 probe on rank 0 while this ctx is still alive, then the ctx is closed: a second live ctx would push
    # the process past CUDA_DEVICE_MAX_CONNECTIONS hardware queues and serialise the pipeline streams of the next one
    extras = {}
    iso = tfl_value = None
    if rank == 0:
        iso = isolated_kernels(g, base)
        tfl = C.c_double(0)
        g._call("debug_fp64_probe", [C.POINTER(C.c_double)], C.byref(tfl))
        tfl_value = float(tfl.value)
    g.close()
    if world > 1 and args.streams != "distinct" and not args.no_extras:
        # SURVEY cfg 5 flavour: every GPU its own scene (distinct maps), a short pass on a fresh ctx
        g2 = capi.Backend(prod, l, r, prm, device=local_rank)
        g2._call("set_pipeline_depth", [C.c_int], args.pipeline_depth)
        base2 = make_workload(seed=edist.stream_seed(rank))
        m2 = measure_stream(args, g2, base2, world, local_rank, ClockSampler(local_rank), K, Wm, prm, want_breakdown=False, target_s=0.1, max_regions=20)
        reg2 = edist.gather_scalars(m2["region_ms"], device="cuda").max(axis=0)
        rec2 = edist.make_record(edist.stream_seed(rank), K, m2["e2e_ctr"], edist.map_checksum(g2.map_download()))
        _, ev2, recs2 = edist.reduce_and_gather(0.0, float(m2["ctr"]["bm_evals"] + m2["lm_exec"]), rec2, device="cuda")
        extras["distinct_streams"] = {"value": ev2 * K / (float(np.median(reg2)) * 1e-3), "unit": "evals/s", "ms_per_step": float(np.median(reg2)) / K,
                                      "note": "every GPU its own scene (seeds 10..): per-frame work differs between GPUs, the step time is the slowest stream's",
                                      "streams": [dict(zip(edist.RECORD_FIELDS, [float(v) for v in q])) for q in recs2]}
        g2.close()
        # BASELINE configs[4]: eight independent 640x480 (DSEC-shaped) streams, one per GPU (here: one per rank that exists)
        try:
            cfg5 = CONFIGS["cfg3"]
            l5, r5 = configs.rig_calibs(cfg5["rig"])
            prm5 = configs.params_for(cfg5["rig"], prod)
            g5 = capi.Backend(prod, l5, r5, prm5, device=local_rank)
            g5._call("set_pipeline_depth", [C.c_int], args.pipeline_depth)
            base5 = make_workload(seed=edist.stream_seed(rank), cfg="cfg3")
            K5 = min(K, 10)
            m5 = measure_stream(args, g5, base5, world, local_rank, ClockSampler(local_rank), K5, 3, prm5, want_breakdown=False, target_s=0.05, max_regions=6)
            reg5 = edist.gather_scalars(m5["region_ms"], device="cuda").max(axis=0)
            reg5e = edist.gather_scalars(m5["e2e_region_ms"], device="cuda").max(axis=0)
            rec5 = edist.make_record(edist.stream_seed(rank), K5, m5["e2e_ctr"], edist.map_checksum(g5.map_download()))
            _, ev5, recs5 = edist.reduce_and_gather(0.0, float(m5["ctr"]["bm_evals"] + m5["lm_exec"]), rec5, device="cuda")
            extras["cfg5_dsec_streams"] = {"workload": f"BASELINE configs[4]: {world} independent 640x480 (DSEC-shaped) stereo event streams, one per GPU (scene seeds 10..)",
                                           "value": ev5 * K5 / (float(np.median(reg5)) * 1e-3), "unit": "evals/s", "ms_per_step": float(np.median(reg5)) / K5,
                                           "e2e_ms_per_step": float(np.median(reg5e)) / K5, "steps_per_region": K5, "regions": int(len(reg5)),
                                           "streams": [dict(zip(edist.RECORD_FIELDS, [float(v) for v in q])) for q in recs5]}
            g5.close()
        except Exception as ex:
            extras["cfg5_dsec_streams"] = {"error": repr(ex)}
    if rank != 0:
        return
    clocks = sampler.stop()
    peak, peak_kind = load_peaks()
    value = evals_all * K / (med_ms / 1e3)
    e2e_value = e2e_evals_all * K / (med_e / 1e3)
    out = {
        "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": med_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 (LM/fusion/tracking), u8/int32 exact (TS, BM moments)", "data": "synthetic",
        "config": {"workload": cfg["name"] + ", one event stream per GPU; step = event ingest + 2 TS builds + BM + LM + cull + "
                               "20-frame window fusion + clean + regularise",
                   "events_into_bm_per_frame": N_SEEDS, "events_per_frame_per_camera": int(base["left"]["x"].size),
                   "parallelism": f"{world} independent streams, one per GPU, no data-path collective",
                   "streams": "every GPU runs the same synthetic stream shape (equal work per GPU)" if args.streams == "same" else "distinct scenes per GPU (seeds 10..)",
                   "l2": "inputs larger than L2, read once: every timed step consumes its own event/seed/pose arrays (1.75 MB per step, "
                         "evicted by a 256 MiB write before the first timed region; with software-pipelined frames in flight a flush "
                         "between iterations would serialise the pipeline); the persistent state (TS images, LUT, grids, ~6 MB) is L2-resident by design",
                   "pipeline_depth": args.pipeline_depth,
                   "priming": "fusion window filled with max_num_fusion_frames frames before the W warm-up steps"},
        "timing": {"regions": int(len(reg)), "steps_per_region": K,
                   "reported": "median region (each region: barrier + sync, K steps, sync + barrier; max over ranks per region)",
                   "timed_s_total": float(np.sum(reg)) / 1e3, "region_ms_min": float(np.min(reg)), "region_ms_max": float(np.max(reg)),
                   "e2e_timed_s_total": float(np.sum(reg_e)) / 1e3, "e2e_region_ms_min": float(np.min(reg_e)), "e2e_region_ms_max": float(np.max(reg_e))},
        "e2e": {"value": e2e_value, "unit": "evals/s", "ms_per_step": med_e / K, "h2d_bytes_per_step": int(m["h2d_per_step"]),
                "d2h_bytes_per_step": int(m["d2h_per_step"]),
                "api": "esvo_stage_ts_events / esvo_run_ts_build / esvo_set_ts_pair_dev / esvo_stage_mapping_inputs / esvo_run_mapping / "
                       "esvo_results_begin / esvo_results_end_view"},
        "gpu_launches": int(m["launches_per_region"]),
        "clocks": clocks,
        "ts_frames_per_s": m["ts_only_frames_per_s"],
        "breakdown_ms_per_step": m["breakdown"],
        "per_step": {"bm_evals": ctr["bm_evals"], "lm_evals_reference_equivalent": ctr["lm_evals"], "lm_evals_executed": lm_exec,
                     "events_into_bm": ctr["n_events"], "n_seeds": ctr["n_seeds"], "n_solved": ctr["n_solved"], "n_culled": ctr["n_culled"],
                     "n_fusions": ctr["n_fusions"], "map_size": ctr["map_size"]},
        "wall_s_timed_region": float(np.sum(reg)) / 1e3,
        "per_rank_ms_per_step": {"resident": [float(v) for v in per_rank[:, 0]], "e2e": [float(v) for v in per_rank[:, 1]],
                                 "host_issue_resident": [float(v) for v in per_rank[:, 2]], "host_issue_e2e": [float(v) for v in per_rank[:, 3]]},
        "host_issue_ms": {"resident": {"mean": float(np.mean(m["issue_resident_ms"]))}, "e2e": {"mean": float(np.mean(m["issue_e2e_ms"]))}},
        "streams": [dict(zip(edist.RECORD_FIELDS, [float(v) for v in q])) for q in records],
    }
    # ---------------- roofline of the dominant kernel: isolated launch, measured FP64 peak ----------------
    ncu = load_ncu_counters()
    lm_ms, bm_ms = iso["lm_ms"], iso["bm_ms"]
    ncand = iso["bm_evals"] / max(iso["n_events"], 1)
    bm_bytes = iso["bm_evals"] * BM_BYTES * (1 + 1 / max(ncand, 1))
    lm_bytes = iso["lm_exec"] * LM_BYTES
    flops_per_eval = ncu.get("fp64_flops_per_executed_eval") if ncu else None
    lm_flops = flops_per_eval * iso["lm_exec"] if flops_per_eval else None
    out["roofline"] = {
        "kernel": "lm2_kernel (DepthProblem LM, one warp per seed)", "bound": "fp64",
        "why": "dependent FP64 issue: the time-surface pair is L1/L2-resident (DRAM traffic ~1.5 % of the algorithmic bytes); ncu shows the FP64 pipe "
               "and the issue slots as the busiest units (profiles/r2_*), so the roofline is the FP64 rate.  The HBM view the north star asked "
               "for is kept under `hbm`.",
        "achieved": (lm_flops / (lm_ms * 1e-3) / 1e12) if lm_flops else None, "peak": tfl_value, "unit": "TFLOP/s",
        "frac": (lm_flops / (lm_ms * 1e-3) / 1e12 / tfl_value) if lm_flops else None,
        "peak_kind": "measured in this run (esvo_debug_fp64_probe: dependent DFMA chains, 2 flops per FMA)",
        "ms_per_launch": lm_ms, "launch": "one isolated launch on an idle GPU (esvo_depth_solve on the frame's seeds), CUDA events inside the library",
        "evals_per_launch": iso["lm_exec"], "seeds_per_launch": iso["n_seeds"], "fp64_flops_per_eval": flops_per_eval,
        "flops_source": (ncu or {}).get("source"), "traffic": (ncu or {}).get("dram_bytes_per_launch"),
        "saturated": (ncu or {}).get("saturated"),
        "hbm": {"bound": "hbm", "achieved": lm_bytes / (lm_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": lm_bytes / (lm_ms * 1e-3) / 1e9 / peak,
                "peak_kind": peak_kind, "algorithmic_bytes_per_launch": lm_bytes,
                "step_level": {"achieved": (ctr["bm_evals"] * BM_BYTES * (1 + 1 / max(ncand, 1)) + lm_exec * LM_BYTES) / (med_ms / K * 1e-3) / 1e9,
                               "note": "BM + LM algorithmic bytes of one step / pipelined ms_per_step"}}}
    out["roofline_other"] = {"kernel": "bm_tma_kernel (EventBM, TMA-staged strip; bm_kernel when ESVO_BM_TMA=0)", "bound": "hbm", "achieved": bm_bytes / (bm_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": bm_bytes / (bm_ms * 1e-3) / 1e9 / peak, "peak_kind": peak_kind, "ms_per_launch": bm_ms,
                             "algorithmic_bytes_per_launch": bm_bytes, "traffic": (ncu or {}).get("bm_dram_bytes_per_launch")}
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_leg(base, sample_steps=3)
        ws = cpu_leg(base, sample_steps=3, shortcut=True)
        out["cpu_baseline"]["with_shortcut"] = {
            "value": ws["value"], "unit": "evals/s", "cores": ws["cores"], "ms_per_step": ws["ms_per_step"], "ts_ms_per_step": ws["ts_ms_per_step"],
            "mapping_ms_per_step": ws["mapping_ms_per_step"],
            "what": "the same port with the CUDA kernel's algorithmic shortcut for the degenerate Student-t scale iteration switched on "
                    "(oracle/o_mapping.h g_irls_shortcut; byte-identical results -- the literal reference loop spins thousands of iterations there)",
            "ratio_value_over_this": value / ws["value"], "ratio_e2e_over_this": e2e_value / ws["value"]}
        out["cpu_baseline"]["ratio_value_over_literal"] = value / out["cpu_baseline"]["value"]
        out["cpu_baseline"]["ratio_e2e_over_literal"] = e2e_value / out["cpu_baseline"]["value"]
        ref4 = cpu_leg(base, sample_steps=1, threads=4)
        out["cpu_baseline"]["reference_thread_config"] = {"value": ref4["value"], "unit": "evals/s", "cores": 4,
                                                           "ms_per_step": ref4["ms_per_step"], "sample": "1 mapping frame, BM+LM on 4 threads"}
    if world == 1 and not args.no_parity:
        par, trk = parity_block(prod, base, cfg["rig"], n_check=args.parity_frames)
        out["parity"] = par
        if trk is not None:
            extras["cfg4_tracking"] = trk
    if world == 1 and not args.no_extras:
        try:
            extras["cfg3"] = run_cfg3(args, prod, local_rank)
        except Exception as ex:       # an extra must never take the headline line down
            extras["cfg3"] = {"error": repr(ex)}
        try:
            extras["sgm_init"] = sgm_block(prod, base, cfg["rig"], local_rank)
        except Exception as ex:
            extras["sgm_init"] = {"error": repr(ex)}
    out["extras"] = extras
    print(json.dumps(out))


def sgm_block(prod, base, rig, local_rank):
    """The one-off initialisation of the mapper (esvo_Mapping::InitializationAtTime, esvo_Mapping.cpp:433-492): cv::StereoSGBM
    (0, 48, 11, 968, 3872, -1, 0, 11) on the time-surface pair on the device (esvo_sgbm_compute, bit-exact vs cv2: tests/test_gpu_sgbm.py)
    followed by esvo_init_from_disparity; host wall clock around the synchronous C-ABI calls, cv2 on the host cores beside it."""
    from esvo_b200 import capi, configs
    l, r = configs.rig_calibs(rig)
    g = capi.Backend(prod, l, r, configs.params_for(rig, prod), device=local_rank)
    for cam, side in ((0, "left"), (1, "right")):
        e = base[side]
        g.ts_push_events(cam, e["x"], e["y"], e["t"], e["p"])
    _, tl = g.ts_build(0, base["t_ts_ns"], want_idx=False); _, tr = g.ts_build(1, base["t_ts_ns"], want_idx=False)
    g.set_ts_pair(None, None, base["T_world_left"])
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); d = g.sgbm_compute(); ts.append((time.perf_counter() - t0) * 1e3)
    sd = base["seeds"]
    t0 = time.perf_counter(); n_pts, acc = g.init_from_disparity(d, sd["x"], sd["y"], base["T_world_left"], 1); t_init = (time.perf_counter() - t0) * 1e3
    out = {"sgbm_ms": float(np.median(ts[1:])), "init_from_disparity_ms": t_init, "sgm_points": int(n_pts), "accepted": bool(acc),
           "valid_disparities": int((d >= 0).sum()), "note": "one-off at start-up; host wall clock around synchronous calls incl. the D2H of the disparity map"}
    g.close()
    try:
        import cv2
        m = cv2.StereoSGBM_create(0, 48, 11, 8 * 121, 32 * 121, -1, 0, 11)
        tc = []
        for _ in range(3):
            t0 = time.perf_counter(); ref = m.compute(tl, tr); tc.append((time.perf_counter() - t0) * 1e3)
        out["cv2_sgbm_ms"] = float(np.median(tc)); out["bit_exact_vs_cv2"] = bool(np.array_equal(ref, d))
    except ImportError:
        pass
    return out


def run_cfg3(args, prod, local_rank):
    """BASELINE configs[2]: 640x480, 20k seeds + fusion (mapping_dsec.yaml).  A short run with the same two legs + parity."""
    from esvo_b200 import capi, configs
    cfg = CONFIGS["cfg3"]
    l, r = configs.rig_calibs(cfg["rig"])
    prm = configs.params_for(cfg["rig"], prod)
    g = capi.Backend(prod, l, r, prm, device=local_rank)
    g._call("set_pipeline_depth", [C.c_int], args.pipeline_depth)
    base = make_workload(seed=3, cfg="cfg3")
    K = min(args.steps, 20)
    m = measure_stream(args, g, base, 1, local_rank, ClockSampler(local_rank), K, 3, prm, want_breakdown=True, target_s=0.15, max_regions=10)
    ctr, lm_exec = m["ctr"], m["lm_exec"]
    med, mede = float(np.median(m["region_ms"])), float(np.median(m["e2e_region_ms"]))
    evals = ctr["bm_evals"] + lm_exec
    res = {"workload": cfg["name"], "value": evals * K / (med * 1e-3), "unit": "evals/s", "ms_per_step": med / K,
           "e2e": {"value": evals * K / (mede * 1e-3), "ms_per_step": mede / K, "h2d_bytes_per_step": int(m["h2d_per_step"]),
                   "d2h_bytes_per_step": int(m["d2h_per_step"])},
           "regions": len(m["region_ms"]), "steps_per_region": K, "per_step": {**{k: ctr[k] for k in CTR_KEYS}, "lm_evals_executed": lm_exec},
           "events_per_frame_per_camera": int(base["left"]["x"].size), "breakdown_ms_per_step": m["breakdown"], "ts_frames_per_s": m["ts_only_frames_per_s"]}
    iso = isolated_kernels(g, base)
    res["isolated_ms"] = {"bm_kernel": iso["bm_ms"], "lm_kernel": iso["lm_ms"], "n_seeds": iso["n_seeds"]}
    g.close()
    if not args.no_parity:
        par, _ = parity_block(prod, base, cfg["rig"], n_check=2, tracking=False)
        res["parity"] = par
    if not args.no_cpu_baseline:
        res["cpu_baseline"] = cpu_leg(base, sample_steps=1, rig=cfg["rig"])
    return res


def _shutdown_pg():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


def cpu_leg(base, sample_steps, threads=None, shortcut=False, rig=RIG):
    """The oracle (a port of the reference CPU path) on the host cores: bounded sample of the same workload."""
    from esvo_b200 import capi, configs
    from oracle.loader import load_oracle
    orc = load_oracle()
    l, r = configs.rig_calibs(rig)
    o = capi.Backend(orc, l, r, configs.params_for(rig, orc))
    nthreads = threads or (os.cpu_count() or 1)
    orc.lib.esvo_oracle_set_exec_threads(o.ctx, C.c_int(nthreads))
    orc.lib.esvo_oracle_set_irls_shortcut(C.c_int(1 if shortcut else 0))
    evals = 0
    ts_s = map_s = 0.0
    t0 = time.perf_counter()
    try:
        for k in range(sample_steps):
            f = shifted(base, k)
            ta = time.perf_counter()
            for cam, side in ((0, "left"), (1, "right")):
                e = f[side]
                o.ts_push_events(cam, e["x"], e["y"], e["t"], e["p"])
                o.ts_build(cam, f["t_ts_ns"], want_idx=False, want_ts=False)
            tb = time.perf_counter()
            o.set_ts_pair(None, None, f["T_world_left"])
            sd = f["seeds"]
            c = o.mapping_at_time(sd["x"], sd["y"], sd["t"], f["pose_t"], f["poses"])
            tc = time.perf_counter()
            ts_s += tb - ta; map_s += tc - tb
            evals += c["bm_evals"] + c["lm_evals"]
    finally:
        orc.lib.esvo_oracle_set_irls_shortcut(C.c_int(0))
    dt = time.perf_counter() - t0
    return {"value": evals / dt, "unit": "evals/s", "cores": nthreads, "kind": "port",
            "sample": f"{sample_steps} mapping frames of the same workload (event ingest + 2 TS builds single-threaded like "
                      f"NUM_THREAD_TS=1; BM+LM on {nthreads} threads with the reference's interleaved fan-out; fusion single-threaded)"
                      + ("; degenerate-IRLS shortcut ON" if shortcut else "; literal reference arithmetic"),
            "ms_per_step": dt / sample_steps * 1e3, "ts_ms_per_step": ts_s / sample_steps * 1e3,
            "mapping_ms_per_step": map_s / sample_steps * 1e3}


def run_reference(args, rank, world):
    if rank != 0:
        return
    base = make_workload(seed=10)
    K, Wm = args.steps, args.warmup
    cpu_leg(base, sample_steps=max(1, min(Wm, 2)))
    leg = cpu_leg(base, sample_steps=K)
    out = {"impl": "reference", "metric": METRIC,
           "value": leg["value"], "unit": "evals/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": leg["ms_per_step"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": CONFIGS["cfg2"]["name"] + ", CPU restatement of the reference (oracle/; the reference itself needs "
                                  "ROS/Eigen/OpenCV C++ and cannot be built here), literal arithmetic, one stream on rank 0"},
           "cpu_baseline": leg,
           "e2e": {"value": leg["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity block (1 GPU)")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg3 / distinct-streams extras")
    ap.add_argument("--parity-frames", type=int, default=3)
    ap.add_argument("--min-timed-s", type=float, default=0.5, help="repeat the K-step region until this much time has been timed")
    ap.add_argument("--streams", default="same", choices=["same", "distinct"], help="per-GPU synthetic streams: same shape (equal work) or distinct scenes")
    ap.add_argument("--scene-seed", type=int, default=10, help="seed of the synthetic scene (rank r uses seed + r with --streams distinct)")
    ap.add_argument("--pipeline-depth", type=int, default=16, help="frames in flight per stream (1 = strictly sequential)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    try:
        if args.impl == "reference":
            run_reference(args, rank, world)
        else:
            run_ours(args, rank, world, local_rank)
    finally:
        _shutdown_pg()


if __name__ == "__main__":
    main()
