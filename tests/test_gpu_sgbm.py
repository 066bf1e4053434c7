"""Device SGBM (esvo_sgbm_compute) against cv2.StereoSGBM, bit for bit, through the C ABI.

The arithmetic is also pinned on the host (tests/test_sgbm_core_host.py runs the very functions the kernels call).  The
device run happens in a subprocess with its own CUDA context (one-off start-up code, kept apart from the hot-path tests).
First hardware run: B200, all three comparisons bit-exact."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from esvo_b200 import capi, configs
from util import scenario, build_ts_pair
prod = capi.load_product()
l, r = configs.rig_calibs("hkust")
g = capi.Backend(prod, l, r, configs.params_for("hkust", prod), device=0)
s = scenario("hkust")
tl, tr = build_ts_pair(g, s)
tl = np.asarray(tl).reshape(260, 346); tr = np.asarray(tr).reshape(260, 346)
d_host = g.sgbm_compute(tl, tr)                      # host images
g.set_ts_pair(tl, tr, s["T_world_left"])
d_dev = g.sgbm_compute()                             # the observation pair that is on the device
d_small = g.sgbm_compute(tl, tr, 32, 5, None, None, 2, 0, 5)
sd = s["seeds"]
n, acc = g.init_from_disparity(d_host, sd["x"], sd["y"], s["T_world_left"], 50)
np.savez(sys.argv[2], tl=tl, tr=tr, d_host=d_host, d_dev=d_dev, d_small=d_small, n=n, acc=acc)
'''


@pytest.mark.gpu
def test_sgbm_device_matches_cv2(tmp_path):
    import cv2
    out = str(tmp_path / "sgbm_out.npz")
    p = subprocess.run([sys.executable, "-c", CHILD, ROOT, out], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    z = np.load(out)
    ref = cv2.StereoSGBM_create(0, 48, 11, 8 * 121, 32 * 121, -1, 0, 11).compute(z["tl"], z["tr"])
    assert (ref >= 0).sum() > 20000
    assert np.array_equal(z["d_host"], ref)
    assert np.array_equal(z["d_dev"], ref)
    ref2 = cv2.StereoSGBM_create(0, 32, 5, 8 * 25, 32 * 25, 2, 0, 5).compute(z["tl"], z["tr"])
    assert np.array_equal(z["d_small"], ref2)
    assert bool(z["acc"]) and int(z["n"]) >= 50
