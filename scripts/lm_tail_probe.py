"""Debug probe: per-seed duration / nfev distribution of the LM kernel on the bench workload."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esvo_b200 import capi, configs, synth
prod = capi.load_product()
s = synth.make_stream("hkust", seed=10, n_seeds=5000, history_ms=50.0)
l, r = configs.rig_calibs("hkust")
g = capi.Backend(prod, l, r, configs.params_for("hkust", prod))
for cam, side in ((0, "left"), (1, "right")):
    e = s[side]; g.ts_push_events(cam, e["x"], e["y"], e["t"], e["p"]); g.ts_build(cam, s["t_ts_ns"])
g.set_ts_pair(None, None, s["T_world_left"])
sd = s["seeds"]
for _ in range(3):
    c = g.mapping_at_time(sd["x"], sd["y"], sd["t"], s["pose_t"], s["poses"])
n = c["n_seeds"]
dbg = np.zeros((n, 4), np.int64)
f = prod.lib.esvo_debug_lm_timing; f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]; f.restype = C.c_int
assert f(g.ctx, dbg.ctypes.data, n) == 0
cyc, nfev, ns, t0 = dbg.T
print("seeds", n, "kernel span us", (t0 + ns).max() / 1e3 - t0.min() / 1e3)
print("per-seed ns: median %.0f p90 %.0f p99 %.0f max %.0f" % tuple(np.percentile(ns, [50, 90, 99, 100])))
print("start offsets us: p50 %.0f p90 %.0f max %.0f" % tuple(np.percentile((t0 - t0.min()) / 1e3, [50, 90, 100])))
order = np.argsort(-ns)[:10]
print("slowest:", [(int(ns[i]), int(nfev[i]), int((t0[i] - t0.min()) / 1e3)) for i in order])
print("nfev hist", np.bincount(nfev.astype(int))[:32])
