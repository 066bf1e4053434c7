"""SURVEY 8f row 4 (esvo_MVStereo comparison modes): the oracle's EventMatcher restatement (oracle/o_event_matcher.h) against an
independent numpy re-derivation written from esvo_core/src/core/EventMatcher.cpp (tests/indep_numpy.py::event_match), and the
EventMatchPair -> DepthPoint conversion (esvo_MVStereo::vEMP2vDP)."""
import numpy as np
import pytest

import indep_numpy as ind
from esvo_b200 import capi, configs
from util import build_ts_pair, em_problem, scenario

RIG = "hkust"


@pytest.fixture(scope="module")
def frame(oracle_lib):
    s = scenario(RIG, seed=4, n_seeds=500)
    l, r = configs.rig_calibs(RIG)
    o = capi.Backend(oracle_lib, l, r, configs.params_for(RIG, oracle_lib))
    tl, tr = build_ts_pair(o, s)
    o.set_ts_pair(tl, tr, s["T_world_left"])
    left, right, counts, poses = em_problem(s)
    return dict(o=o, s=s, tl=tl, tr=tr, left=left, right=right, counts=counts, poses=poses)


def test_event_slicing_covers_the_window(frame):
    c = frame["counts"]
    assert c.size >= 3 and (c > 0).all() and c.sum() <= frame["left"]["t"].size
    # slices are contiguous: every slice but the last spans at least its thickness minus one event gap
    t = frame["left"]["t"]; b = np.concatenate([[0], np.cumsum(c)])
    for k in range(c.size - 1):
        assert t[b[k + 1] - 1] - t[b[k]] >= 0 and t[b[k + 1]] > t[b[k + 1] - 1] - 1


@pytest.mark.parametrize("kw", [dict(time_thr=5e-4, epi_thr=1.0, ncc_thr=0.1, patch=(15, 7), num_thread=4),
                                dict(time_thr=1e-4, epi_thr=0.5, ncc_thr=0.3, patch=(25, 5), num_thread=3)])
def test_event_matcher_vs_numpy(frame, kw):
    f = frame; o = f["o"]
    seeds, evals = o.em_match(f["left"], f["right"], f["counts"], f["poses"], **kw)
    cams = configs.rig_arrays(RIG)
    _, _, lut_l, _ = o.get_rectify_tables(0); _, _, lut_r, _ = o.get_rectify_tables(1)
    ref, evals_ref = ind.event_match(f["left"], f["right"], f["counts"], f["poses"], lut_l, lut_r, cams["left"]["P"], cams["right"]["P"],
                                     o.get_derived()["baseline"], f["tl"], f["tr"], f["s"]["T_world_left"], kw["time_thr"], kw["epi_thr"],
                                     kw["ncc_thr"], kw["patch"][0], kw["patch"][1], kw["num_thread"])
    assert len(ref) > 50, len(ref)
    assert seeds.size == len(ref) and evals == evals_ref
    for sd, m in zip(seeds, ref):
        assert sd["t_ns"] == m["t_ns"]
        assert np.array_equal(sd["x_left"], m["x_left"]) and np.array_equal(sd["x_right"], m["x_right"])
        assert abs(sd["cost"] - m["cost"]) < 1e-12
        assert abs(sd["inv_depth"] - m["inv_depth"]) <= 1e-15 * abs(m["inv_depth"]) * 4
        assert np.array_equal(sd["T_world_virtual"].reshape(4, 4), m["T"])
    # the matcher recovers the scene: most matched depths agree with the block matcher's range
    inv = seeds["inv_depth"]
    assert ((inv > 0.2) & (inv < 2.5)).mean() > 0.7


def test_vemp_to_points_vs_numpy(frame):
    f = frame; o = f["o"]
    seeds, _ = o.em_match(f["left"], f["right"], f["counts"], f["poses"])
    pts = o.seeds_to_points(seeds)
    Pl = configs.rig_arrays(RIG)["left"]["P"]
    age = int(configs.params_for(RIG, o.L).age_vis_threshold)
    for sd, p in zip(seeds[:200], pts[:200]):
        m = dict(x_left=sd["x_left"], inv_depth=sd["inv_depth"], cost=sd["cost"], T=sd["T_world_virtual"].reshape(4, 4))
        q = ind.vemp_to_points([m], Pl, age)[0]
        assert p["row"] == q["row"] and p["col"] == q["col"] and np.array_equal(p["x"], q["x"])
        assert p["inv_depth"] == q["inv_depth"] and p["variance"] == 1e-6 and p["residual"] == q["residual"] and p["age"] == q["age"]
        assert np.allclose(p["p_cam"], q["p_cam"], rtol=1e-12, atol=0)
        assert np.array_equal(p["T_world_cam"].reshape(4, 4), q["T"])


def test_naive_propagation_of_matched_points(frame):
    """Mode PURE_EVENT_MATCHING downstream of the matcher (esvo_MVStereo.cpp:272-286): two vectors of matched points (the second
    from a shifted pose) splat into the frame, newest first -- the oracle against the independent Python re-derivation of
    DepthFusion::naive_propagation (tests/indep_fusion.py): element order, coordinates, values."""
    import indep_fusion as inf
    f = frame; o = f["o"]
    seeds, _ = o.em_match(f["left"], f["right"], f["counts"], f["poses"])
    pts = o.seeds_to_points(seeds)
    older = pts.copy()
    older["T_world_cam"][:, 3] += 0.004; older["T_world_cam"][:, 11] -= 0.003
    older["residual"] *= np.where(np.arange(older.size) % 3 == 0, 0.2, 1.0)      # some of them win the replacement test
    Pl = configs.rig_arrays(RIG)["left"]["P"]
    T = f["s"]["T_world_left"]
    grid = inf.Grid(o.H, o.W)
    for k, v in enumerate((pts, older)):
        o.naive_propagate(v, T, k == 0)
        vec = [dict(p_cam=q["p_cam"], var=float(q["variance"]), res=float(q["residual"]), age=int(q["age"]), T_world_cam=q["T_world_cam"]) for q in v]
        inf.naive_propagate_vector(grid, vec, T, Pl, o.W, o.H)
        m = o.map_download()
        assert len(grid.elements) == m.size and m.size > 100
        assert [e.row for e in grid.elements] == m["row"].tolist() and [e.col for e in grid.elements] == m["col"].tolist()
        for name, get in (("inv_depth", lambda e: e.rho), ("variance", lambda e: e.var), ("residual", lambda e: e.res)):
            assert np.allclose(np.array([get(e) for e in grid.elements]), m[name], rtol=1e-10, atol=0), (k, name)
        assert [e.age for e in grid.elements] == m["age"].tolist()
        assert np.allclose(np.array([e.p_cam for e in grid.elements]), m["p_cam"], rtol=1e-9, atol=1e-12)
    assert any((e.row, e.col) != k for k, e in grid.cell.items()), "no replacement happened"
