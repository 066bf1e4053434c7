"""Shared helpers for the parity tests: one synthetic scenario, two backends fed identical inputs."""
import functools

import numpy as np

from esvo_b200 import capi, configs, synth


@functools.lru_cache(maxsize=8)
def scenario(rig="hkust", seed=2, n_seeds=1500, t_ts=0.5):
    kw = {}
    if rig == "dsec":
        kw = dict(n_segments=120)
    return synth.make_stream(rig, seed=seed, n_seeds=n_seeds, t_ts=t_ts, **kw)


def make_backends(rig, oracle_lib, product_lib, tweak=None):
    l, r = configs.rig_calibs(rig)
    po = configs.params_for(rig, oracle_lib)
    pp = configs.params_for(rig, product_lib)
    if tweak:
        tweak(po); tweak(pp)
    o = capi.Backend(oracle_lib, l, r, po)
    g = capi.Backend(product_lib, l, r, pp)
    # The product runs on ITS OWN rectification tables (host_setup.cpp, pinned to cv2 bit for bit in
    # tests/test_product_tables_cv2.py); the checker gets the same tables so that both sides see identical inputs.
    for cam in (0, 1):
        m1, m2, lut, mask = g.get_rectify_tables(cam)
        o.set_rectify_tables(cam, m1, m2, lut, mask)
    return o, g


def build_ts_pair(b, s):
    for cam, side in ((0, "left"), (1, "right")):
        e = s[side]
        b.ts_push_events(cam, e["x"], e["y"], e["t"], e["p"])
    _, tl = b.ts_build(0, s["t_ts_ns"], want_idx=False)
    _, tr = b.ts_build(1, s["t_ts_ns"], want_idx=False)
    return tl, tr


def rel(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def em_problem(s, window_ms=4.0, slice_thickness=1e-3, max_events=3000):
    """The inputs esvo_MVStereo hands its EventMatcher (esvo_MVStereo.cpp:579-608,1008-1040): the left / right events of the
    window before the observation stamp (at most EM_NUM_EVENT_MATCHING + 1 each), the left ones cut into slices of
    EM_Slice_Thickness with the pose at each slice's median stamp."""
    import indep_numpy as ind
    t_up = s["t_ts_ns"]; t_low = t_up - int(window_ms * 1e6)

    def window(e):
        lo = int(np.searchsorted(e["t"], t_low, side="left")); hi = int(np.searchsorted(e["t"], t_up, side="left")) - 1
        hi = min(hi, lo + max_events + 1)
        return {k: e[k][lo:hi].copy() for k in ("x", "y", "t", "p")}
    left, right = window(s["left"]), window(s["right"])
    counts, med = ind.event_slicing_for_em(left["t"], t_low, t_up, slice_thickness)
    poses = []
    for t in med:
        T = synth.pose_at(t * 1e-9)
        T[:3, 3] *= 1.0 if s["rig"] == "hkust" else 20.0     # synth.make_stream's scene scale
        poses.append(T.ravel())
    return left, right, counts, np.array(poses)
