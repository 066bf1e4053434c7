"""Debug probe: per-seed duration / evaluations / IRLS trips of the LM kernel on the bench workload (one isolated launch)."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esvo_b200 import capi, configs, synth
prod = capi.load_product()
s = synth.make_stream("hkust", seed=int(os.environ.get("PROBE_SEED", "10")), n_seeds=5000, history_ms=50.0)
l, r = configs.rig_calibs("hkust")
g = capi.Backend(prod, l, r, configs.params_for("hkust", prod))
for cam, side in ((0, "left"), (1, "right")):
    e = s[side]; g.ts_push_events(cam, e["x"], e["y"], e["t"], e["p"]); g.ts_build(cam, s["t_ts_ns"], want_idx=False, want_ts=False)
g.set_ts_pair(None, None, s["T_world_left"])
sd = s["seeds"]
seeds, _ = g.bm_match(sd["x"], sd["y"], sd["t"], s["pose_t"], s["poses"])
for _ in range(3):
    pts, nfev = g.depth_solve(seeds)
n = seeds.size
dbg = np.zeros((n, 4), np.int64)
f = prod.lib.esvo_debug_lm_timing; f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]; f.restype = C.c_int
assert f(g.ctx, dbg.ctypes.data, n) == 0
cyc, nfev_s, nexec, trips = dbg.T
us = cyc / 1965.0
print("seeds", n, "per-seed us: mean %.0f median %.0f p90 %.0f p99 %.0f max %.0f" % (us.mean(), *np.percentile(us, [50, 90, 99, 100])))
print("nexec: mean %.1f max %d; trips: mean %.0f p90 %.0f p99 %.0f max %d; trips per pair-eval mean %.1f" % (nexec.mean(), nexec.max(), trips.mean(), *np.percentile(trips, [90, 99]), trips.max(), trips.sum() / (nexec.sum() / 2)))
o = np.argsort(-us)[:12]
print("slowest:", [(int(us[i]), int(nfev_s[i]), int(nexec[i]), int(trips[i])) for i in o])
print("corr(us, trips) %.3f corr(us, nexec) %.3f" % (np.corrcoef(us, trips)[0, 1], np.corrcoef(us, nexec)[0, 1]))
print("us per trip (fit): %.3f ; us per exec-eval beyond trips: %.2f" % tuple(np.linalg.lstsq(np.stack([trips, nexec], 1).astype(float), us, rcond=None)[0]))
