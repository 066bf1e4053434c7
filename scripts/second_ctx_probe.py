import sys, os, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from esvo_b200 import capi, configs
class A: pass
args = A(); args.pipeline_depth = 16
prod = capi.load_product()
l, r = configs.rig_calibs("hkust"); prm = configs.params_for("hkust", prod)
torch.cuda.set_device(0)
for trial, seed, bd in ((0, 10, True), (1, 10, False), (2, 11, False), (3, 10, True)):
    g = capi.Backend(prod, l, r, prm, device=0)
    g._call("set_pipeline_depth", [C.c_int], 16)
    base = bench.make_workload(seed=seed)
    m = bench.measure_stream(args, g, base, 1, 0, bench.ClockSampler(0), 20, 5, prm, want_breakdown=bd, target_s=0.1, max_regions=20)
    print("trial", trial, "seed", seed, "breakdown", bd, "resident ms/step %.4f e2e %.4f regions %d" % (np.median(m["region_ms"]) / 20, np.median(m["e2e_region_ms"]) / 20, len(m["region_ms"])), flush=True)
    g.close()
