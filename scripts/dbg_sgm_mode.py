"""debug: MVStereo SGM mode, product vs oracle -- which map pixels differ and which points touch them"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cv2
from esvo_b200 import capi, configs
from oracle.loader import load_oracle
from util import build_ts_pair, make_backends, scenario
from test_gpu_mvstereo import _frame, _run_mode
orc = load_oracle(); prod = capi.load_product()
o = g = None
wo, wg = [], []
for k, t_ts in enumerate((0.50, 0.53, 0.56)):
    s, o, g = _frame(orc, prod, "hkust", t_ts=t_ts, backends=(o, g) if o else None)
    s = dict(s); s["Pl00"] = configs.rig_arrays("hkust")["left"]["P"][0, 0]
    d_dev = g.sgbm_compute()
    mo = _run_mode(o, 4, s, wo, 2, d_dev); mg = _run_mode(g, 4, s, wg, 2, d_dev)
    print("frame", k, "window equal:", [a.tobytes() == b.tobytes() for a, b in zip(wo, wg)], "map", mo.size, mg.size)
    so = {(int(a["row"]), int(a["col"])) for a in mo}; sg = {(int(a["row"]), int(a["col"])) for a in mg}
    print(" only oracle:", sorted(so - sg)[:12], " only product:", sorted(sg - so)[:12])
    # re-project every window point on the host and list those near the differing pixels
    T = np.asarray(s["T_world_left"], float); Tfw = np.linalg.inv(T); P = configs.rig_arrays("hkust")["left"]["P"]
    for (r, c) in sorted((so ^ sg))[:6]:
        for wi, v in enumerate(wo):
            pc = v["p_cam"]; Tw = v["T_world_cam"].reshape(-1, 4, 4)
            pw = np.einsum("nij,nj->ni", Tw[:, :3, :3], pc) + Tw[:, :3, 3]
            pf = pw @ Tfw[:3, :3].T + Tfw[:3, 3]
            h = pf @ P[:, :3].T + P[:, 3]
            x = h[:, 0] / h[:, 2]; y = h[:, 1] / h[:, 2]
            near = np.nonzero((np.abs(x - c) < 1.5) & (np.abs(y - r) < 1.5))[0]
            for q in near[:4]:
                print("  pixel", (r, c), "vector", wi, "point", q, "x,y = %.15f %.15f" % (x[q], y[q]), "rho", v["inv_depth"][q])
