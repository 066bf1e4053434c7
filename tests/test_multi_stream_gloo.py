"""world_size-2 CPU test (gloo) of the N>1 path: one independent stream per rank, no data-path collective,
timing = max over ranks, work = sum over ranks, one result record per stream gathered on every rank.
The per-rank compute is done by the oracle here (CPU); on the GPU box bench.py runs the same plumbing over NCCL."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from esvo_b200 import capi, configs, dist as edist, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.loader import load_oracle
    lib = load_oracle()
    seed = edist.stream_seed(rank)
    s = synth.make_stream("hkust", seed=seed, n_seeds=200, n_segments=40, history_ms=30.0)
    l, r = configs.rig_calibs("hkust")
    o = capi.Backend(lib, l, r, configs.params_for("hkust", lib))
    for cam, side in ((0, "left"), (1, "right")):
        e = s[side]
        o.ts_push_events(cam, e["x"], e["y"], e["t"], e["p"])
        o.ts_build(cam, s["t_ts_ns"])
    o.set_ts_pair(None, None, s["T_world_left"])
    sd = s["seeds"]
    c = o.mapping_at_time(sd["x"], sd["y"], sd["t"], s["pose_t"], s["poses"])
    rec = edist.make_record(seed, 1, c, edist.map_checksum(o.map_download()))
    local_ms = 10.0 + 5.0 * rank
    mx, tot, recs = edist.reduce_and_gather(local_ms, c["bm_evals"] + c["lm_evals"], rec)
    per_rank = edist.gather_scalars([local_ms, float(rank)])          # bench.py's per-rank diagnostics
    assert per_rank.shape == (world, 2) and list(per_rank[:, 0]) == [10.0, 15.0] and list(per_rank[:, 1]) == [0.0, 1.0]
    q.put((rank, mx, tot, recs, c["bm_evals"] + c["lm_evals"]))
    dist.destroy_process_group()


def test_two_streams_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, mx0, tot0, recs0, ev0), (r1, mx1, tot1, recs1, ev1) = res
    assert mx0 == mx1 == 15.0                      # max over ranks
    assert tot0 == tot1 == ev0 + ev1               # whole-job work
    assert np.array_equal(recs0, recs1) and recs0.shape == (2, 10)
    assert list(recs0[:, 0]) == [10.0, 11.0]       # one stream per rank, seeds 10 and 11
    assert recs0[0, 9] != recs0[1, 9]              # different streams -> different maps
    assert recs0[0, 6] + recs0[0, 7] == ev0 and recs0[1, 6] + recs0[1, 7] == ev1


def test_single_rank_fallback():
    from esvo_b200 import dist as edist
    rec = np.arange(10, dtype=np.float64)
    mx, tot, recs = edist.reduce_and_gather(3.0, 7.0, rec)
    assert (mx, tot) == (3.0, 7.0) and recs.shape == (1, 10)
