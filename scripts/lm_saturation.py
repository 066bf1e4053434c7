"""LM kernel alone and saturated: the bench workload's seeds (one frame, ~2.3 k) and the same seeds tiled 16x
(~37 k seeds = 15 waves, no tail), timed with CUDA events around the kernel (esvo_profile stage 3).
usage: python scripts/lm_saturation.py [variant ...]   (ESVO_LM_VARIANT codes, one subprocess each)"""
import ctypes as C, json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(tag):
    from esvo_b200 import capi, configs, synth
    prod = capi.load_product()
    s = synth.make_stream("hkust", seed=10, n_seeds=5000, history_ms=50.0)
    l, r = configs.rig_calibs("hkust")
    g = capi.Backend(prod, l, r, configs.params_for("hkust", prod))
    for cam, side in ((0, "left"), (1, "right")):
        e = s[side]; g.ts_push_events(cam, e["x"], e["y"], e["t"], e["p"]); g.ts_build(cam, s["t_ts_ns"])
    g.set_ts_pair(None, None, s["T_world_left"])
    sd = s["seeds"]
    seeds, _ = g.bm_match(sd["x"], sd["y"], sd["t"], s["pose_t"], s["poses"])
    f64, u64 = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
    out = {"variant": tag, "n_seeds": int(seeds.size)}
    for name, rep in (("one_frame", 1), ("x16", 16)):
        sd_t = np.tile(seeds, rep)
        g._call("profile", [C.c_int], 1 << 3)
        g._call("profile_read", [f64, u64], (C.c_double * 8)(), (C.c_uint64 * 8)())
        ts = []
        for it in range(4):
            pts, ev = g.depth_solve(sd_t)
            ms = (C.c_double * 8)(); cnt = (C.c_uint64 * 8)()
            g._call("profile_read", [f64, u64], ms, cnt)
            ts.append(ms[3])
        out[name] = {"ms": min(ts[1:]), "all": ts, "ms_per_frame": min(ts[1:]) / rep, "n_pts": int(pts.size), "nfev": int(ev)}
        if rep == 1:
            np.save(os.path.join(ROOT, "gpurun_out", f"lm_rho_{tag}.npy"), pts["inv_depth"])
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        variants = sys.argv[1:] or ["1600", "1601", "1611", "2001", "2011", "2401", "2411"]
        for v in variants:
            env = dict(os.environ, ESVO_LM_VARIANT=v)
            subprocess.run([sys.executable, __file__, "--child", v], env=env)
        base = None
        for v in variants:
            p = os.path.join(ROOT, "gpurun_out", f"lm_rho_{v}.npy")
            if not os.path.exists(p):
                continue
            a = np.load(p)
            if base is None:
                base = a
            else:
                rel = np.abs(a - base) / np.abs(base) if a.shape == base.shape else None
                print(v, "vs", variants[0], "shape", a.shape, "max rel", None if rel is None else float(rel.max()),
                      "median", None if rel is None else float(np.median(rel)))
