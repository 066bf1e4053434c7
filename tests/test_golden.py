"""Golden vectors (tests/golden/small_rig_frame.npz, produced by scripts/make_golden.py from the oracle).
CPU: the oracle still reproduces them bit for bit.  GPU: the CUDA path reproduces them (integers exactly,
f64 within 1e-4 relative)."""
import ast
import os

import numpy as np
import pytest

from esvo_b200 import capi, configs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "small_rig_frame.npz")


def _load():
    z = np.load(GOLD, allow_pickle=False)
    rig = ast.literal_eval(bytes(z["rig_json"]).decode())
    configs.RIGS["golden_small"] = rig
    return z


def _backend(lib, z):
    l, r = configs.rig_calibs("golden_small")
    p = configs.params_for("hkust", lib)
    p.max_num_fusion_frames = 2
    b = capi.Backend(lib, l, r, p)
    b.set_rectify_tables(0, z["map1_l"], z["map2_l"], z["lut_l"], z["mask_l"])
    b.set_rectify_tables(1, z["map1_r"], z["map2_r"], z["lut_r"], z["mask_r"])
    return b


def _run(b, z, exact):
    out = {}
    for cam, side in ((0, "left"), (1, "right")):
        ev, t = z[f"ev_{side}"], z[f"evt_{side}"]
        b.ts_push_events(cam, ev[0], ev[1], t, ev[2].astype(np.uint8))
    idx_l, ts_l = b.ts_build(0, int(z["t_ts_ns"]))
    _, ts_r = b.ts_build(1, int(z["t_ts_ns"]))
    idx_m, ts_m = b.ts_build(0, int(z["t_mid_ns"]))
    assert np.array_equal(idx_l, z["idx_left"]) and np.array_equal(idx_m, z["idx_mid"])
    assert np.array_equal(ts_l, z["ts_left"]) and np.array_equal(ts_r, z["ts_right"]) and np.array_equal(ts_m, z["ts_mid"])
    b.set_ts_pair(ts_l, ts_r, z["T_world_left"])
    sx, st = z["seeds_xy"], z["seeds_t"]
    seeds, ev = b.bm_match(sx[0], sx[1], st, z["pose_t"], z["poses"])
    g = z["bm_seeds"]
    assert ev == int(z["bm_evals"]) and seeds.size == g.size
    assert np.array_equal(seeds["x_left"], g["x_left"]) and np.array_equal(seeds["disp"], g["disp"])
    assert np.allclose(seeds["cost"], g["cost"], rtol=0, atol=1e-12)
    pts, _ = b.depth_solve(g)
    gp = z["lm_points"]
    assert pts.size == gp.size
    r = np.abs(pts["inv_depth"] - gp["inv_depth"]) / np.abs(gp["inv_depth"])
    assert (r < 1e-4).mean() > 0.995 and np.median(r) < 1e-7
    if exact:
        assert pts.tobytes() == gp.tobytes() and seeds.tobytes() == g.tobytes()
    for key in ("frame1", "frame2"):
        c = b.mapping_at_time(sx[0], sx[1], st, z["pose_t"], z["poses"])
        gc = z[f"{key}_counters"]
        got = np.array(list(c.values()), np.int64)
        assert np.array_equal(got[[0, 1, 2, 3, 4, 5, 7]], gc[[0, 1, 2, 3, 4, 5, 7]]), (key, c, gc)
        m, gm = b.map_download(), z[f"{key}_map"]
        assert m.size == gm.size and np.array_equal(m["row"], gm["row"]) and np.array_equal(m["col"], gm["col"])
        rr = np.abs(m["inv_depth"] - gm["inv_depth"]) / np.abs(gm["inv_depth"])
        assert (rr < 1e-4).mean() > 0.995
        if exact:
            assert m.tobytes() == gm.tobytes()
    cloud = z["trk_cloud"].copy()
    b.track_srand(1)
    assert b.track_reset(cloud, z["T_world_left"], z["trk_prior"], ts_l) == 0
    T, stt = b.track_solve(True)
    assert [stt["n_points"], stt["nfev"], stt["n_iter"]] == list(z["trk_stats"])
    assert np.abs(T - z["trk_pose"]).max() < 1e-6
    if exact:
        assert np.array_equal(T, z["trk_pose"])
    return out


def test_oracle_reproduces_golden(oracle_lib):
    z = _load()
    _run(_backend(oracle_lib, z), z, exact=True)


@pytest.mark.gpu
def test_cuda_path_reproduces_golden(product_lib):
    z = _load()
    _run(_backend(product_lib, z), z, exact=False)


def test_oracle_reproduces_golden_extras(oracle_lib):
    """tests/golden/extras.npz (scripts/make_golden.py extras): FORWARD time surface, out-of-order stamps, initialisation
    from an SGM disparity map -- the oracle still reproduces them bit for bit."""
    z = _load()
    x = np.load(os.path.join(os.path.dirname(GOLD), "extras.npz"), allow_pickle=False)
    l, r = configs.rig_calibs("golden_small")

    def backend(tweak=None):
        p = configs.params_for("hkust", oracle_lib)
        if tweak:
            tweak(p)
        b = capi.Backend(oracle_lib, l, r, p)
        b.set_rectify_tables(0, z["map1_l"], z["map2_l"], z["lut_l"], z["mask_l"])
        b.set_rectify_tables(1, z["map1_r"], z["map2_r"], z["lut_r"], z["mask_r"])
        return b

    ev, t = z["ev_left"], z["evt_left"]
    def fw(p):
        p.time_surface_mode = 1; p.ignore_polarity = 0
    o = backend(fw)
    o.ts_push_events(0, ev[0], ev[1], t, ev[2].astype(np.uint8))
    assert np.array_equal(o.ts_build(0, int(z["t_ts_ns"]))[1], x["fwd_T"])
    assert np.array_equal(o.ts_build(0, int(z["t_mid_ns"]))[1], x["fwd_mid"])
    o = backend()
    o.ts_push_events(0, ev[0], ev[1], x["t_jitter"], ev[2].astype(np.uint8))
    i1, s1 = o.ts_build(0, int(z["t_ts_ns"])); i2, s2 = o.ts_build(0, int(z["t_mid_ns"]))
    assert np.array_equal(i1, x["un_idx"]) and np.array_equal(s1, x["un_ts"])
    assert np.array_equal(i2, x["un_idx_mid"]) and np.array_equal(s2, x["un_ts_mid"])
    assert (x["un_idx"] != z["idx_left"]).any()            # the jitter does change which arrival a pixel reports
    o = backend()
    n, acc = o.init_from_disparity(x["disp16"], z["seeds_xy"][0], z["seeds_xy"][1], z["T_world_left"], 20)
    assert acc and n == int(x["sgm_n"])
    assert o.window_download(0).tobytes() == x["sgm_points"].tobytes()
    assert o.map_download().tobytes() == x["sgm_map"].tobytes()
