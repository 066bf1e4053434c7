#!/usr/bin/env python
"""Generates tests/golden/*.npz: seeded synthetic inputs + the outputs of the CPU oracle.

The reference has no tests or golden vectors and cannot be built in this image (ROS / Eigen / OpenCV C++
are absent), so these vectors are produced by oracle/ (the line-by-line CPU restatement, whose third-party
pieces are pinned against cv2 / MINPACK / libc in tests/test_oracle_*.py).  Re-run after a deliberate
change of the oracle:   python scripts/make_golden.py
Inputs are small (a 120 x 96 crop-sized rig derived from the hkust calibration) to keep the fixtures < 1 MB.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from esvo_b200 import capi, configs, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def small_rig():
    """hkust intrinsics scaled to 120 x 96 (keeps distortion, rectification and the stereo baseline)."""
    r = configs.RIGS["hkust"]
    sx, sy = 120 / 346, 96 / 260
    out = dict(width=120, height=96, model="plumb_bob")
    for side in ("left", "right"):
        c = r[side]
        K = np.array(c["K"], float).reshape(3, 3).copy(); K[0] *= sx; K[1] *= sy
        P = np.array(c["P"], float).reshape(3, 4).copy(); P[0] *= sx; P[1] *= sy
        if side == "right":
            K[2] = [0, 0, 1]
        out[side] = dict(K=K.ravel().tolist(), D=c["D"], R=c["R"], P=P.ravel().tolist())
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    configs.RIGS["golden_small"] = small_rig()
    from oracle.loader import load_oracle
    lib = load_oracle()
    l, r = configs.rig_calibs("golden_small")
    p = configs.params_for("hkust", lib)
    p.max_num_fusion_frames = 2
    p.invdepth_min_range, p.invdepth_max_range = 0.25, 2.0
    o = capi.Backend(lib, l, r, p)
    # scene
    import esvo_b200.synth as S
    s = S.make_stream("golden_small", seed=3, n_seeds=400, n_segments=25, depth_range=(0.6, 3.0), history_ms=40.0)
    ts, idx = [], []
    for cam, side in ((0, "left"), (1, "right")):
        e = s[side]
        o.ts_push_events(cam, e["x"], e["y"], e["t"], e["p"])
        i, t = o.ts_build(cam, s["t_ts_ns"])
        ts.append(t); idx.append(i)
    T_mid = int(s["left"]["t"][s["left"]["t"].size // 2])
    idx_mid, ts_mid = o.ts_build(0, T_mid)
    o.set_ts_pair(ts[0], ts[1], s["T_world_left"])
    sd = s["seeds"]
    seeds, bm_evals = o.bm_match(sd["x"], sd["y"], sd["t"], s["pose_t"], s["poses"])
    pts, lm_evals = o.depth_solve(seeds)
    ctr = o.mapping_at_time(sd["x"], sd["y"], sd["t"], s["pose_t"], s["poses"])
    m1 = o.map_download()
    ctr2 = o.mapping_at_time(sd["x"], sd["y"], sd["t"], s["pose_t"], s["poses"])
    m2 = o.map_download()
    # tracking
    pts3 = s["scene_points"]
    rng = np.random.default_rng(9)
    cloud = pts3[rng.choice(pts3.shape[0], 1500, replace=False)].astype(np.float32)
    Tp = s["T_world_left"].copy(); Tp[:3, 3] += [0.003, -0.002, 0.001]
    c_in = cloud.copy()
    p2 = o.params
    o.track_srand(1)
    rc = o.track_reset(c_in, s["T_world_left"], Tp, ts[0])
    assert rc == 0
    T_trk, st = o.track_solve(True)
    tables = [o.get_rectify_tables(c) for c in (0, 1)]
    np.savez_compressed(
        os.path.join(OUT, "small_rig_frame.npz"),
        rig_json=np.frombuffer(repr(configs.RIGS["golden_small"]).encode(), np.uint8),
        ev_left=np.stack([s["left"]["x"], s["left"]["y"], s["left"]["p"]]).astype(np.uint16), evt_left=s["left"]["t"],
        ev_right=np.stack([s["right"]["x"], s["right"]["y"], s["right"]["p"]]).astype(np.uint16), evt_right=s["right"]["t"],
        t_ts_ns=np.int64(s["t_ts_ns"]), t_mid_ns=np.int64(T_mid), T_world_left=s["T_world_left"],
        seeds_xy=np.stack([sd["x"], sd["y"]]).astype(np.uint16), seeds_t=sd["t"], pose_t=s["pose_t"], poses=s["poses"],
        map1_l=tables[0][0], map2_l=tables[0][1],
        map1_r=tables[1][0], map2_r=tables[1][1], lut_l=tables[0][2], mask_l=tables[0][3], lut_r=tables[1][2], mask_r=tables[1][3],
        ts_left=ts[0], ts_right=ts[1], idx_left=idx[0].astype(np.int32), idx_mid=idx_mid.astype(np.int32), ts_mid=ts_mid,
        bm_seeds=seeds, bm_evals=np.int64(bm_evals), lm_points=pts, lm_evals=np.int64(lm_evals),
        frame1_counters=np.array(list(ctr.values()), np.int64), frame1_map=m1,
        frame2_counters=np.array(list(ctr2.values()), np.int64), frame2_map=m2,
        trk_cloud=cloud, trk_prior=Tp, trk_pose=T_trk, trk_stats=np.array([st["n_points"], st["nfev"], st["n_iter"]], np.int64),
    )
    print("wrote", os.path.join(OUT, "small_rig_frame.npz"), os.path.getsize(os.path.join(OUT, "small_rig_frame.npz")), "bytes")
    print("bm seeds", seeds.size, "lm points", pts.size, "map", m1.size, m2.size, "track", st)


def extras():
    """tests/golden/extras.npz: outputs of the oracle for the paths added after the first fixture -- FORWARD time surface,
    out-of-order stamps, initialisation from an SGM disparity map -- on the inputs of small_rig_frame.npz."""
    import ast
    import cv2
    z = np.load(os.path.join(OUT, "small_rig_frame.npz"), allow_pickle=False)
    configs.RIGS["golden_small"] = ast.literal_eval(bytes(z["rig_json"]).decode())
    from oracle.loader import load_oracle
    lib = load_oracle()
    l, r = configs.rig_calibs("golden_small")

    def backend(tweak=None):
        p = configs.params_for("hkust", lib)
        if tweak:
            tweak(p)
        b = capi.Backend(lib, l, r, p)
        b.set_rectify_tables(0, z["map1_l"], z["map2_l"], z["lut_l"], z["mask_l"])
        b.set_rectify_tables(1, z["map1_r"], z["map2_r"], z["lut_r"], z["mask_r"])
        return b

    ev, t = z["ev_left"], z["evt_left"]
    # FORWARD mode, polarity on, median 3x3
    def fw(p):
        p.time_surface_mode = 1; p.ignore_polarity = 0
    o = backend(fw)
    o.ts_push_events(0, ev[0], ev[1], t, ev[2].astype(np.uint8))
    _, fwd_T = o.ts_build(0, int(z["t_ts_ns"]))
    _, fwd_mid = o.ts_build(0, int(z["t_mid_ns"]))
    # out-of-order stamps: every 11th event arrives 0.3 ms "late"
    tj = t.copy(); tj[5::11] -= 300_000
    o = backend()
    o.ts_push_events(0, ev[0], ev[1], tj, ev[2].astype(np.uint8))
    un_idx, un_ts = o.ts_build(0, int(z["t_ts_ns"]))
    un_idx_mid, un_ts_mid = o.ts_build(0, int(z["t_mid_ns"]))
    # initialisation from the SGM disparity of the golden TS pair (cv::StereoSGBM with the reference's parameters)
    H, W = z["ts_left"].shape
    disp = cv2.StereoSGBM_create(0, 48, 11, 8 * 11 * 11, 32 * 11 * 11, -1, 0, 11).compute(z["ts_left"], z["ts_right"])
    o = backend()
    n, acc = o.init_from_disparity(disp, z["seeds_xy"][0], z["seeds_xy"][1], z["T_world_left"], 20)
    assert acc, n
    np.savez_compressed(os.path.join(OUT, "extras.npz"), fwd_T=fwd_T, fwd_mid=fwd_mid, t_jitter=tj, un_idx=un_idx.astype(np.int32), un_ts=un_ts,
                        un_idx_mid=un_idx_mid.astype(np.int32), un_ts_mid=un_ts_mid, disp16=disp.astype(np.int16), sgm_n=np.int64(n),
                        sgm_points=o.window_download(0), sgm_map=o.map_download())
    print("wrote", os.path.join(OUT, "extras.npz"), os.path.getsize(os.path.join(OUT, "extras.npz")), "bytes; sgm points", n)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "extras":
        extras()
    else:
        main()
