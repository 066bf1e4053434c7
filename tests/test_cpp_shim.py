"""The reference-shaped C++ shim (include/esvo_b200/esvo_core.hpp) compiles against the C ABI with plain g++,
fails loudly without a GPU (CPU box) and runs the BM -> LM -> cull -> fuse chain on the GPU box."""
import os
import subprocess

import pytest

from conftest import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "esvo_b200", "_build", "example_mapping_frame")


def _build():
    build = os.path.join(ROOT, "esvo_b200", "_build")
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "mapping_frame.cpp"),
                           "-L" + build, "-lesvo_b200", "-Wl,-rpath," + build, "-o", EXE])


def test_cpp_shim_compiles_and_fails_loudly_without_gpu(product_lib):
    _build()
    if has_gpu():
        return
    p = subprocess.run([EXE], capture_output=True, text=True)
    assert p.returncode == 2 and "no CPU fallback" in p.stdout


@pytest.mark.gpu
def test_cpp_shim_runs_on_gpu(product_lib):
    _build()
    p = subprocess.run([EXE], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "BM 500 seeds" in p.stdout, p.stdout


def test_cpp_shim_event_frontend_helpers(tmp_path):
    """esvo_core::frontend (selectCloseEvents / samplePoseStamps, esvo_Mapping.cpp:536-603) -- host logic, header-only."""
    exe = str(tmp_path / "frontend_check")
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "frontend_check.cpp"), "-o", exe])
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr


def test_pipelined_c_abi_example_compiles(product_lib, tmp_path):
    """examples/pipelined_stream.cpp (plain C ABI, S frames in flight): builds with g++; without a GPU it stops at esvo_create."""
    build = os.path.join(ROOT, "esvo_b200", "_build")
    exe = str(tmp_path / "pipelined_stream")
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "pipelined_stream.cpp"),
                           "-L" + build, "-lesvo_b200", "-Wl,-rpath," + build, "-o", exe])
    if has_gpu():
        return
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 2 and "no CPU fallback" in p.stdout


@pytest.mark.gpu
def test_shim_mapping_to_tracking_matches_oracle(oracle_lib, product_lib, tmp_path):
    """examples/shim_loop.cpp: dataTransferring helpers -> esvo_Mapping::MappingAtTime -> packPointCloud -> esvo_Tracking
    (refMapCallback / timeSurfaceCallback / eventsCallback / TrackingLoopOnce) through the C++ shim on the GPU, against the same
    sequence on the oracle with the front-end restated in Python from the reference (tests/indep_numpy.py)."""
    import struct
    import numpy as np
    import indep_numpy as ind
    from esvo_b200 import capi, configs, synth
    build = os.path.join(ROOT, "esvo_b200", "_build")
    exe = str(tmp_path / "shim_loop")
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "shim_loop.cpp"),
                           "-L" + build, "-lesvo_b200", "-Wl,-rpath," + build, "-o", exe])
    s = synth.make_stream("hkust", seed=2, n_seeds=5000, t_ts=0.5)
    s2 = synth.make_stream("hkust", seed=2, n_seeds=100, t_ts=0.51)
    l, r = configs.rig_calibs("hkust")
    o = capi.Backend(oracle_lib, l, r, configs.params_for("hkust", oracle_lib))
    for cam, side in ((0, "left"), (1, "right")):
        e = s[side]; o.ts_push_events(cam, e["x"], e["y"], e["t"], e["p"])
    _, tl = o.ts_build(0, s["t_ts_ns"], want_idx=False); _, tr = o.ts_build(1, s["t_ts_ns"], want_idx=False)
    o.ts_reset(0)
    e2 = s2["left"]; o.ts_push_events(0, e2["x"], e2["y"], e2["t"], e2["p"])
    _, tc = o.ts_build(0, s2["t_ts_ns"], want_idx=False)
    W, H = o.W, o.H
    half_slice, pen = 0.001, 5000
    L = s["left"]
    Tw = np.ascontiguousarray(s["T_world_left"], np.float64)
    scen = tmp_path / "scen.bin"; resf = tmp_path / "res.bin"
    with open(scen, "wb") as f:
        f.write(struct.pack("<iii", W, H, 0)); f.write(struct.pack("<qq", s["t_ts_ns"], s2["t_ts_ns"])); f.write(Tw.tobytes())
        f.write(struct.pack("<d", half_slice)); f.write(struct.pack("<ii", pen, L["x"].size))
        for a, dt in ((L["x"], np.uint16), (L["y"], np.uint16), (L["t"], np.int64), (L["p"], np.uint8)):
            f.write(np.ascontiguousarray(a, dt).tobytes())
        f.write(tl.tobytes()); f.write(tr.tobytes()); f.write(tc.tobytes())
        f.write(struct.pack("<i", s["pose_t"].size)); f.write(np.ascontiguousarray(s["pose_t"], np.int64).tobytes())
        f.write(np.ascontiguousarray(s["poses"], np.float64).tobytes())
    p = subprocess.run([exe, str(scen), str(resf)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    raw = open(resf, "rb").read()
    hdr = np.frombuffer(raw, np.int32, 8); off = 32
    ctr = np.frombuffer(raw, np.uint64, 8, off); off += 64
    T_idle = np.frombuffer(raw, np.float64, 16, off).reshape(4, 4); off += 128
    T_work = np.frombuffer(raw, np.float64, 16, off).reshape(4, 4); off += 128
    sel = np.frombuffer(raw, np.int64, hdr[0], off); off += 8 * hdr[0]
    stamps = np.frombuffer(raw, np.int64, hdr[1], off); off += 8 * hdr[1]
    cloud = np.frombuffer(raw, np.float32, 3 * hdr[3], off).reshape(-1, 3)
    # ---- the same sequence on the oracle, front-end restated from the reference ----
    sel_ref = ind.select_close_events(L["t"], s["t_ts_ns"], half_slice, pen)
    assert np.array_equal(sel, sel_ref) and sel.size == pen - 1          # observation stamp newer than every event: one budget slot skipped
    st_ref = ind.sample_pose_stamps(s["t_ts_ns"], half_slice)
    idx = np.searchsorted(s["pose_t"], st_ref, side="left")
    keep = idx < s["pose_t"].size
    assert np.array_equal(stamps, st_ref[keep])
    poses = s["poses"][idx[keep]]
    o.set_ts_pair(tl, tr, Tw)
    co = o.mapping_at_time(L["x"][sel], L["y"][sel], L["t"][sel], st_ref[keep], poses)
    assert [int(v) for v in ctr] == [co[k] for k in ("n_events", "n_seeds", "n_solved", "n_culled", "n_fusions", "bm_evals", "lm_evals", "map_size")][:8] or \
        [int(v) for v in ctr[:6]] == [co[k] for k in ("n_events", "n_seeds", "n_solved", "n_culled", "n_fusions", "bm_evals")]
    mo = o.map_download()
    assert hdr[2] == mo.size and hdr[3] == mo.size
    cloud_ref = ind.pack_point_cloud(mo["p_cam"], Tw)
    rel = np.linalg.norm(cloud - cloud_ref, axis=1) / np.linalg.norm(cloud_ref, axis=1)
    assert (rel < 1e-5).mean() > 0.999, rel.max()
    # tracking, node already WORKING: reference pose = pose at the map stamp, prior = last tracked pose (= T here)
    assert hdr[5] == 1
    c = cloud.copy(); o.track_srand(1)
    assert o.track_reset(c, Tw, Tw, tc) == 0
    To, st = o.track_solve(True)
    assert hdr[7] == st["n_iter"]
    assert np.abs(T_work - To).max() < 1e-6, np.abs(T_work - To).max()
    # node in INITIALIZATION / IDLE: reference pose = identity, prior = reference pose (esvo_Tracking.cpp:183-184,228-233)
    assert hdr[4] == 1
    c = cloud.copy(); o.track_srand(1)
    o.track_reset(c, np.eye(4), np.eye(4), tc)
    To2, _ = o.track_solve(True)
    assert np.abs(T_idle - To2).max() < 1e-6
    # numEventsSinceLastObs_ = distance(lower_bound(old cur stamp = 0), lower_bound(t_cur)) + 1 (:241-243)
    assert hdr[6] == int(np.searchsorted(L["t"], s2["t_ts_ns"], side="left")) + 1


@pytest.mark.gpu
def test_shim_mvstereo_modes_match_oracle(oracle_lib, product_lib, tmp_path):
    """examples/mvstereo_modes.cpp: esvo_core::esvo_MVStereo::MappingAtTime in its five MVStereoMode settings (event matcher [26],
    block matching, EM + optimisation, BM + optimisation = the ESVO mapper, SGM) through the C++ shim on the GPU, one frame each,
    against the same sequences on the oracle (tests/test_gpu_mvstereo.py::_run_mode; cv2.StereoSGBM feeds the oracle's SGM mode)."""
    import struct
    import numpy as np
    cv2 = pytest.importorskip("cv2")
    from esvo_b200 import capi, configs
    from test_gpu_mvstereo import _run_mode
    from util import build_ts_pair, em_problem, make_backends, scenario
    build = os.path.join(ROOT, "esvo_b200", "_build")
    exe = str(tmp_path / "mvstereo_modes")
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "mvstereo_modes.cpp"),
                           "-L" + build, "-lesvo_b200", "-Wl,-rpath," + build, "-o", exe])
    s = dict(scenario("hkust", seed=4, n_seeds=800, t_ts=0.5))
    s["Pl00"] = configs.rig_arrays("hkust")["left"]["P"][0, 0]
    o, g = make_backends("hkust", oracle_lib, product_lib)
    g.close()
    tl, tr = build_ts_pair(o, s)
    o.ts_reset(0); o.ts_reset(1)
    left, right, counts, poses = em_problem(s)
    t_up = s["t_ts_ns"]; t_low = t_up - int(4.0 * 1e6)
    import indep_numpy as ind
    _, med = ind.event_slicing_for_em(left["t"], t_low, t_up, 1e-3)
    Tw = np.ascontiguousarray(s["T_world_left"], np.float64)
    sd = s["seeds"]
    scen = tmp_path / "scen.bin"; resf = tmp_path / "res.bin"
    with open(scen, "wb") as f:
        f.write(struct.pack("<ii", o.W, o.H)); f.write(struct.pack("<qqq", s["t_ts_ns"], t_low, t_up)); f.write(Tw.tobytes())
        f.write(tl.tobytes()); f.write(tr.tobytes())
        for e in (left, right):
            f.write(struct.pack("<i", e["x"].size))
            for a, dt in ((e["x"], np.uint16), (e["y"], np.uint16), (e["t"], np.int64), (e["p"], np.uint8)):
                f.write(np.ascontiguousarray(a, dt).tobytes())
        f.write(struct.pack("<i", sd["x"].size))
        for a, dt in ((sd["x"], np.uint16), (sd["y"], np.uint16), (sd["t"], np.int64)):
            f.write(np.ascontiguousarray(a, dt).tobytes())
        f.write(struct.pack("<i", s["pose_t"].size)); f.write(np.ascontiguousarray(s["pose_t"], np.int64).tobytes())
        f.write(np.ascontiguousarray(s["poses"], np.float64).tobytes())
        f.write(struct.pack("<i", med.size)); f.write(np.ascontiguousarray(med, np.int64).tobytes())
        f.write(np.ascontiguousarray(poses, np.float64).tobytes())
    p = subprocess.run([exe, str(scen), str(resf)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    raw = open(resf, "rb").read()
    rec = np.dtype([("row", "<i4"), ("col", "<i4"), ("inv_depth", "<f8"), ("variance", "<f8"), ("residual", "<f8"), ("age", "<i8")])
    off = 0
    d_ref = cv2.StereoSGBM_create(0, 48, 11, 8 * 121, 32 * 121, -1, 0, 11).compute(tl, tr)
    for mode in range(5):
        n_match, n_map = struct.unpack_from("<ii", raw, off); off += 8
        m = np.frombuffer(raw, rec, n_map, off); off += n_map * rec.itemsize
        o.set_ts_pair(tl, tr, Tw)
        mo = _run_mode(o, mode, s, [], 20, d_ref)
        assert n_map == mo.size and n_map > 20, (mode, n_map, mo.size, p.stdout)
        assert np.array_equal(m["row"], mo["row"]) and np.array_equal(m["col"], mo["col"]) and np.array_equal(m["age"], mo["age"]), mode
        r = np.abs(m["inv_depth"] - mo["inv_depth"]) / np.abs(mo["inv_depth"])
        assert (r < 1e-4).mean() > 0.995 and np.median(r) < 1e-7, (mode, np.median(r), r.max())


def _nccl_build(exe):
    build = os.path.join(ROOT, "esvo_b200", "_build")
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
                           os.path.join(ROOT, "examples", "multi_stream_nccl.cpp"), "-L" + build, "-lesvo_b200", "-Wl,-rpath," + build,
                           "-L/usr/local/cuda/lib64", "-lcudart", "-lnccl", "-lpthread", "-o", exe])


def test_multi_stream_nccl_example_compiles(product_lib, tmp_path):
    """examples/multi_stream_nccl.cpp (C++ host, one thread per GPU, ncclAllGather of the stream records): builds against the
    system NCCL; without a GPU it stops before any work."""
    if not os.path.exists("/usr/include/nccl.h"):
        pytest.skip("no system NCCL headers")
    exe = str(tmp_path / "multi_stream_nccl")
    _nccl_build(exe)
    if has_gpu():
        return
    p = subprocess.run([exe, "/dev/null"], capture_output=True, text=True)
    assert p.returncode == 2 and "no CPU fallback" in p.stdout


@pytest.mark.gpu
def test_multi_stream_nccl_gather_matches_python_records(oracle_lib, product_lib, tmp_path):
    """The C++ multi-stream host: every visible GPU runs its own stream (rank r drops 200 r events from the selection budget),
    the records travel through ONE ncclAllGather; each record equals what the Python binding computes for the same stream
    (esvo_b200/dist.py layout and checksum)."""
    import struct
    import numpy as np
    import indep_numpy as ind
    from esvo_b200 import capi, configs, dist as edist, synth
    if not os.path.exists("/usr/include/nccl.h"):
        pytest.skip("no system NCCL headers")
    exe = str(tmp_path / "multi_stream_nccl")
    _nccl_build(exe)
    s = synth.make_stream("hkust", seed=2, n_seeds=5000, t_ts=0.5)
    l, r = configs.rig_calibs("hkust")
    o = capi.Backend(oracle_lib, l, r, configs.params_for("hkust", oracle_lib))
    for cam, side in ((0, "left"), (1, "right")):
        e = s[side]; o.ts_push_events(cam, e["x"], e["y"], e["t"], e["p"])
    _, tl = o.ts_build(0, s["t_ts_ns"], want_idx=False); _, tr = o.ts_build(1, s["t_ts_ns"], want_idx=False)
    W, H = o.W, o.H
    half_slice, pen, frames = 0.001, 3000, 2
    L = s["left"]
    Tw = np.ascontiguousarray(s["T_world_left"], np.float64)
    scen = tmp_path / "scen.bin"
    with open(scen, "wb") as f:
        f.write(struct.pack("<iii", W, H, 0)); f.write(struct.pack("<qq", s["t_ts_ns"], s["t_ts_ns"])); f.write(Tw.tobytes())
        f.write(struct.pack("<d", half_slice)); f.write(struct.pack("<ii", pen, L["x"].size))
        for a, dt in ((L["x"], np.uint16), (L["y"], np.uint16), (L["t"], np.int64), (L["p"], np.uint8)):
            f.write(np.ascontiguousarray(a, dt).tobytes())
        f.write(tl.tobytes()); f.write(tr.tobytes()); f.write(tl.tobytes())
        f.write(struct.pack("<i", s["pose_t"].size)); f.write(np.ascontiguousarray(s["pose_t"], np.int64).tobytes())
        f.write(np.ascontiguousarray(s["poses"], np.float64).tobytes())
    p = subprocess.run([exe, str(scen), "8", str(frames)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = p.stdout.strip().splitlines()                      # NCCL may print its version banner on stdout first
    world = int([ln for ln in lines if ln.startswith("streams ")][0].split()[1])
    recs = np.array([[float(v) for v in ln.split()[1:]] for ln in lines if ln.startswith("record ")])
    assert recs.shape == (world, len(edist.RECORD_FIELDS)) and world >= 1
    print("C++ multi-stream host:", world, "stream(s)")
    st_ref = ind.sample_pose_stamps(s["t_ts_ns"], half_slice)
    idx = np.searchsorted(s["pose_t"], st_ref, side="left"); keep = idx < s["pose_t"].size
    for rank in range(world):
        g = capi.Backend(product_lib, l, r, configs.params_for("hkust", product_lib))
        sel = ind.select_close_events(L["t"], s["t_ts_ns"], half_slice, pen - 200 * rank)
        g.set_ts_pair(tl, tr, Tw)
        for _ in range(frames):
            c = g.mapping_at_time(L["x"][sel], L["y"][sel], L["t"][sel], st_ref[keep], s["poses"][idx[keep]])
        rec = edist.make_record(rank, frames, c, edist.map_checksum(g.map_download()))
        g.close()
        assert np.array_equal(recs[rank][:7], rec[:7]) and recs[rank][8] == rec[8], (rank, recs[rank], rec)
        assert abs(recs[rank][7] - rec[7]) <= 0.01 * rec[7]                                  # lm_evals (nfev-style counter)
        assert abs(recs[rank][9] - rec[9]) <= 1e-9 * abs(rec[9]), (rank, recs[rank][9], rec[9])


@pytest.mark.parametrize("epoch_ns", [0, 1600000000 * 10**9])
def test_frontend_selections_vs_numpy(tmp_path, epoch_ns):
    """examples/frontend_dump.cpp: what esvo_core::frontend selects from an event buffer (selectCloseEvents, selectSGMEvents,
    samplePoseStamps, eventSlicingForEM) against the independent numpy re-derivation of esvo_Mapping::dataTransferring /
    esvo_MVStereo::eventSlicingForEM (tests/indep_numpy.py) -- also with UNIX-epoch stamps, where ros::Time::toSec() collapses
    neighbouring nanosecond stamps into one double and the lower bounds compare those doubles."""
    import struct
    import numpy as np
    import indep_numpy as ind
    exe = str(tmp_path / "frontend_dump")
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "frontend_dump.cpp"), "-o", exe])
    rng = np.random.default_rng(3)
    n = 40000
    t = epoch_ns + np.cumsum(rng.integers(1, 900, n)).astype(np.int64)          # ~0.45 us apart: several stamps per double at epoch scale
    for t_end, half_slice, pen in ((int(t[n // 2]), 0.001, 5000), (int(t[-1]) + 1000, 0.001, 3000), (int(t[n // 3]) + 137, 0.0005, 100000)):
        t_up = t_end; t_low = t_end - 4_000_000
        inp = tmp_path / "in.bin"; outp = tmp_path / "out.bin"
        with open(inp, "wb") as f:
            f.write(struct.pack("<i", n)); f.write(t.tobytes()); f.write(struct.pack("<qdiqqd", t_end, half_slice, pen, t_low, t_up, 1e-3))
        p = subprocess.run([exe, str(inp), str(outp)], capture_output=True, text=True)
        assert p.returncode == 0, p.stdout + p.stderr
        raw = open(outp, "rb").read(); off = 0
        def arr(dtype=np.int64):
            nonlocal off
            m = struct.unpack_from("<i", raw, off)[0]; off += 4
            a = np.frombuffer(raw, dtype, m, off); off += m * np.dtype(dtype).itemsize
            return a
        close, sgm, stamps = arr(), arr(), arr()
        sl = arr(np.dtype([("count", "<i4"), ("median", "<i8")]))
        assert np.array_equal(close, ind.select_close_events(t, t_end, half_slice, pen))
        assert np.array_equal(sgm, ind.select_sgm_events(t, t_end, half_slice, pen))
        assert np.array_equal(stamps, ind.sample_pose_stamps(t_end, half_slice))
        w = t[(t >= t_low) & (t < t_up)][:-1]
        counts, med = ind.event_slicing_for_em(w, t_low, t_up, 1e-3)
        assert np.array_equal(sl["count"], counts) and np.array_equal(sl["median"], med)
        assert close.size > 0 and stamps.size > 100 and counts.size >= 3
