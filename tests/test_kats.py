"""Closed-form known-answer tests: anchors that do not depend on the oracle being right.
Run against the oracle on CPU and against the CUDA path on the GPU box."""
import ctypes as C

import numpy as np
import pytest

from esvo_b200 import capi, configs
from conftest import has_gpu


def _backends():
    out = [pytest.param("oracle", id="oracle")]
    out.append(pytest.param("cuda", id="cuda", marks=pytest.mark.gpu))
    return out


def _make(kind, oracle_lib, request, tweak=None, identity_maps=True):
    lib = oracle_lib if kind == "oracle" else request.getfixturevalue("product_lib")
    l, r = configs.rig_calibs("hkust")
    p = configs.params_for("hkust", lib)
    if tweak:
        tweak(p)
    b = capi.Backend(lib, l, r, p)
    if identity_maps:
        H, W = b.H, b.W
        yy, xx = np.mgrid[0:H, 0:W]
        lut = np.stack([xx, yy], -1).astype(np.float64)
        for cam in (0, 1):
            b.set_rectify_tables(cam, xx.astype(np.float32), yy.astype(np.float32), lut, np.full((H, W), 255, np.uint8))
    return b


@pytest.mark.parametrize("kind", _backends())
def test_kat_time_surface_decay(kind, oracle_lib, request):
    """3x3 event blob at t_e, sync at T: centre = rne(255 exp(-(T - t_e)/0.030)) (TimeSurface.cpp:75-77,126-127)."""
    b = _make(kind, oracle_lib, request)
    t_e = 2_000_000_000
    xs, ys = np.meshgrid([100, 101, 102], [50, 51, 52])
    n = xs.size
    b.ts_push_events(0, xs.ravel(), ys.ravel(), np.full(n, t_e, np.int64), np.ones(n, np.uint8))
    for dt_ms in (0.001, 5.0, 12.345, 30.0, 90.0):
        T = t_e + int(dt_ms * 1e6)
        idx, ts = b.ts_build(0, T)
        want = int(np.rint(255.0 * np.exp(-(dt_ms * 1e-3) / 0.030)))
        assert ts[51, 101] == want, (dt_ms, ts[51, 101], want)
        assert idx[51, 101] == 4 and idx[0, 0] == -1 and ts[0, 0] == 0


@pytest.mark.parametrize("kind", _backends())
def test_kat_td_stdvar_and_disparity_clip(kind, oracle_lib, request):
    b = _make(kind, oracle_lib, request)
    d = b.get_derived()
    assert abs(d["td_stdvar"] - 56.5347) < 2e-3          # cfg/mapping/mapping_rpg.yaml:26-28
    assert (d["min_disparity"], d["max_disparity"]) == (3, 28)   # hkust: f*b = 13.8634, rho in [0.25, 2]
    assert abs(d["baseline"] - 13.8634 / 189.705) < 1e-9


@pytest.mark.parametrize("kind", _backends())
def test_kat_fronto_parallel_disparity(kind, oracle_lib, request):
    """Right TS = left TS shifted by an integer disparity d: BM returns d, LM returns rho ~= d/(f b)
    (EventBM.cpp:147-152)."""
    b = _make(kind, oracle_lib, request)
    H, W, d = b.H, b.W, 9
    rng = np.random.default_rng(0)
    left = np.zeros((H, W), np.uint8)
    tex = rng.integers(30, 255, (H, W)).astype(np.uint8)
    left[40:220, 60:300] = tex[40:220, 60:300]
    right = np.zeros_like(left)
    right[:, : W - d] = left[:, d:]
    T = np.eye(4)
    b.set_ts_pair(left, right, T)
    ex = rng.integers(100, 250, 200).astype(np.uint16); ey = rng.integers(60, 200, 200).astype(np.uint16)
    et = np.full(200, 1_000_000_000, np.int64)
    seeds, _ = b.bm_match(ex, ey, et, np.array([1_000_000_000], np.int64), T.reshape(1, 16))
    assert seeds.size == 200
    assert np.all(seeds["disp"] == d) and np.all(seeds["cost"] < 1e-7)   # not exactly 0: sigma carries +1e-6
    fb = 13.8634
    assert np.allclose(seeds["inv_depth"], d / fb, rtol=1e-12)
    pts, _ = b.depth_solve(seeds)
    assert pts.size == 200
    assert np.allclose(pts["inv_depth"], d / fb, rtol=2e-3)
    assert np.all(pts["residual"] < 1e-3)


@pytest.mark.parametrize("kind", _backends())
def test_kat_student_t_fusion(kind, oracle_lib, request):
    """Fusing two identical estimates (rho, s2, nu): rho, s2*nu/(2(nu+1)), nu+1; age incremented twice
    (DepthPoint.cpp:171-179 + DepthFusion.cpp:171)."""
    b = _make(kind, oracle_lib, request)
    pt = np.zeros(1, capi.DEPTH_POINT_DTYPE)
    rho, s2, nu = 0.8, 1e-4, 2.1897
    pt["row"], pt["col"] = 100, 150
    pt["x"] = [150.25, 100.75]
    pt["inv_depth"], pt["scale2"], pt["nu"] = rho, s2, nu
    pt["variance"] = nu / (nu - 2) * s2
    pt["residual"] = 5.0
    fx, cx, cy = 189.705, 165.382, 121.295
    z = 1 / rho
    pt["p_cam"] = [(150.25 - cx) * z / fx, (100.75 - cy) * z / fx, z]
    pt["T_world_cam"] = np.eye(4).ravel()
    assert b.fuse(pt, np.eye(4), 0, True) == 0
    m = b.map_download()
    assert m.size == 4 and sorted(zip(m["row"], m["col"])) == [(100, 150), (100, 151), (101, 150), (101, 151)]
    assert np.allclose(m["inv_depth"], rho) and np.allclose(m["scale2"], s2) and np.all(m["age"] == 0)
    assert b.fuse(pt, np.eye(4), 0, False) == 4
    m = b.map_download()
    assert np.allclose(m["inv_depth"], rho, rtol=1e-12)
    assert np.allclose(m["scale2"], s2 * nu / (2 * (nu + 1)), rtol=1e-12)
    assert np.allclose(m["nu"], nu + 1) and np.all(m["age"] == 2)
    # incompatible and farther estimate: skipped by the occlusion rule (DepthFusion.cpp:181)
    far = pt.copy(); far["inv_depth"] = 0.2; far["p_cam"] = pt["p_cam"] * (rho / 0.2)
    assert b.fuse(far, np.eye(4), 0, False) == 0
    m2 = b.map_download()
    assert np.allclose(m2["inv_depth"], rho, rtol=1e-12)


@pytest.mark.parametrize("kind", _backends())
def test_kat_tracking_zero_motion(kind, oracle_lib, request):
    """Perfect map, zero relative motion: the pose stays put (RegProblemLM.cpp:348-372)."""
    b = _make(kind, oracle_lib, request)
    H, W = b.H, b.W
    rng = np.random.default_rng(4)
    fx, cx, cy = 189.705, 165.382, 121.295
    n = 1500
    u = rng.uniform(30, W - 30, n); v = rng.uniform(30, H - 30, n); z = rng.uniform(0.8, 3.0, n)
    ts = np.zeros((H, W), np.uint8)
    ts[np.round(v).astype(int), np.round(u).astype(int)] = 255       # events exactly where the map projects
    cloud = np.stack([(np.round(u) - cx) * z / fx, (np.round(v) - cy) * z / fx, z], 1).astype(np.float32)
    b.track_srand(1)
    assert b.track_reset(cloud, np.eye(4), np.eye(4), ts) == 0
    T, st = b.track_solve(True)
    assert st["n_points"] == 500 and st["n_iter"] >= 1
    assert np.abs(T - np.eye(4)).max() < 2e-3


def test_kat_zncc_cost(oracle_lib):
    """ZNCC(p,p) ~= 0, ZNCC(p, c - p) ~= 1; sigma carries +1e-6 (utils.h:74-92, EventBM.cpp:317-333)."""
    rng = np.random.default_rng(1)
    p = rng.integers(0, 256, 105).astype(np.float64)
    f = oracle_lib.lib.esvo_oracle_op_zncc
    f.restype = C.c_double
    ptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert abs(f(ptr(p), ptr(p), C.c_size_t(105))) < 1e-7
    q = 255.0 - p
    assert abs(f(ptr(p), ptr(q), C.c_size_t(105)) - 1.0) < 1e-7
    flat = np.full(105, 7.0)
    assert abs(f(ptr(flat), ptr(p), C.c_size_t(105)) - 0.5) < 1e-12   # zero variance -> correlation 0


def test_oracle_forward_time_surface_against_numpy(oracle_lib):
    """FORWARD mode of createTimeSurfaceAtTime (TimeSurface.cpp:86-116) restated independently in numpy:
    raster-order bilinear splat of exp(-dt/tau) with a clamp at 1 after every accumulation."""
    from util import scenario
    l, r = configs.rig_calibs("hkust")
    p = configs.params_for("hkust", oracle_lib)
    p.time_surface_mode = 1; p.median_blur_kernel_size = 0
    o = capi.Backend(oracle_lib, l, r, p)
    s = scenario("hkust"); e = s["left"]
    o.ts_push_events(0, e["x"], e["y"], e["t"], e["p"])
    T = s["t_ts_ns"]
    idx, img = o.ts_build(0, T)
    _, _, lut, _ = o.get_rectify_tables(0)
    W, H = 346, 260
    idx = np.asarray(idx).reshape(-1); lut = np.asarray(lut).reshape(-1, 2)
    mp = np.zeros((H, W))
    for pix in np.nonzero(idx >= 0)[0]:                      # ascending = raster order
        ev = np.exp(-((T - int(e["t"][idx[pix]])) * 1e-9) / 0.03)
        u, v = lut[pix]
        if not (u >= 0 and v >= 0):
            continue
        ui, vi = int(np.floor(u)), int(np.floor(v))
        if not (ui + 1 < W and vi + 1 < H):
            continue
        fu, fv = u - ui, v - vi
        for yy, xx, w in ((vi, ui, (1 - fu) * (1 - fv)), (vi, ui + 1, fu * (1 - fv)), (vi + 1, ui, (1 - fu) * fv), (vi + 1, ui + 1, fu * fv)):
            mp[yy, xx] = min(mp[yy, xx] + w * ev, 1.0)
    ref = np.clip(np.rint(255 * mp), 0, 255).astype(np.uint8)
    img = np.asarray(img).reshape(H, W)
    assert (img == 255).sum() > 100                           # the clamp is exercised
    assert np.abs(ref.astype(int) - img.astype(int)).max() <= 1 and (ref != img).mean() < 1e-3


def test_oracle_init_from_disparity_kat(oracle_lib):
    """InitializationAtTime after SGM (esvo_Mapping.cpp:446-486): rho = disp / (P00 * baseline); events on pixels with
    invalid or out-of-range disparity are dropped; naive_propagation keeps the NEAREST point per pixel."""
    l, r = configs.rig_calibs("hkust")
    p = configs.params_for("hkust", oracle_lib)
    o = capi.Backend(oracle_lib, l, r, p)
    W, H = 346, 260
    fb = o.get_derived()["baseline"] * 189.705
    disp = np.full((H, W), -16, np.int16)                 # invalid everywhere ...
    disp[100:140, 100:200] = 16 * 10                      # ... except a block at 10 px  -> rho = 10 / fb
    disp[120:125, 150:160] = 16 * 20                      # and a nearer patch at 20 px
    disp[10:20, 10:20] = 16 * 1                           # rho below invDepth_min_range (0.25): 1 / 13.86 = 0.072 -> dropped
    _, _, lut, _ = o.get_rectify_tables(0)
    lut = np.asarray(lut).reshape(H, W, 2)
    ys, xs = np.mgrid[90:150:3, 90:210:3]
    ex = xs.ravel().astype(np.uint16); ey = ys.ravel().astype(np.uint16)
    T = np.eye(4)
    n, acc = o.init_from_disparity(disp, ex, ey, T, min_points=10)
    # expectation from the LUT
    cx = np.floor(lut[ey, ex, 0]).astype(int); cy = np.floor(lut[ey, ex, 1]).astype(int)
    inside = (cx >= 0) & (cx < W) & (cy >= 0) & (cy < H)
    d = np.where(inside, disp[np.clip(cy, 0, H - 1), np.clip(cx, 0, W - 1)], -16) / 16.0
    rho = d / fb
    keep = inside & (d >= 0) & (rho >= 0.25) & (rho <= 2.0)
    assert acc and n == keep.sum() and n > 100
    m = o.map_download()
    assert m.size > n                                      # 2x2 splats
    got = {(int(e["row"]), int(e["col"])): float(e["inv_depth"]) for e in m}
    assert set(np.round(list(got.values()), 12)) <= set(np.round(np.unique(rho[keep]), 12))
    assert abs(max(got.values()) - 20 / fb) < 1e-12 and abs(min(got.values()) - 10 / fb) < 1e-12
    n2, acc2 = o.init_from_disparity(disp, ex[:5], ey[:5], T, min_points=10)
    assert not acc2 and n2 <= 5


def test_oracle_unordered_stamps_quirk(oracle_lib):
    """eventsCallback (TimeSurface.cpp:410-422) queues events_.back() -- the latest STAMP so far -- not the new event:
    an event that arrives with an older stamp re-queues the current latest event at ITS pixel instead."""
    l, r = configs.rig_calibs("hkust")
    o = capi.Backend(oracle_lib, l, r, configs.params_for("hkust", oracle_lib))
    x = np.array([10, 20, 30, 40], np.uint16); y = np.array([5, 5, 5, 5], np.uint16)
    t = np.array([10, 30, 20, 40], np.int64) + 1_000_000_000; p = np.ones(4, np.uint8)
    o.ts_push_events(0, x, y, t, p)
    idx, _ = o.ts_build(0, int(t.max()) + 1)
    idx = np.asarray(idx).reshape(260, 346)
    assert idx[5, 10] == 0 and idx[5, 20] == 2 and idx[5, 30] == -1 and idx[5, 40] == 3
    # T between the stamps: the entry queued by arrival 2 carries stamp 30
    idx2, _ = o.ts_build(0, 1_000_000_025)
    idx2 = np.asarray(idx2).reshape(260, 346)
    assert idx2[5, 10] == 0 and idx2[5, 20] == -1 and idx2[5, 40] == -1
