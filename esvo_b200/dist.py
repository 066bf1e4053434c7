"""Multi-GPU plumbing: independent stereo event streams are sharded one per GPU.  There is no
collective inside the hot path; torch.distributed (NCCL on GPUs, gloo in CPU tests) is used only to agree on
the timing (max over ranks), to add up the work done (sum over ranks) and to gather one fixed-size result
record per stream on rank 0."""
from __future__ import annotations

import numpy as np

RECORD_FIELDS = ("stream_id", "frames", "n_seeds", "n_solved", "n_culled", "n_fusions", "bm_evals", "lm_evals",
                 "map_size", "map_checksum")


def stream_seed(rank: int, base: int = 10) -> int:
    """Seed of the synthetic stream handled by `rank` (BASELINE configs[4]: seeds 10..17 for 8 GPUs)."""
    return base + rank


def map_checksum(map_points: np.ndarray) -> float:
    """Order-sensitive checksum of a downloaded map: sum_i (i+1) * (row_i*4096 + col_i + rho_i)."""
    if map_points.size == 0:
        return 0.0
    i = np.arange(1, map_points.size + 1, dtype=np.float64)
    return float(np.sum(i * (map_points["row"] * 4096.0 + map_points["col"] + map_points["inv_depth"])))


def make_record(stream_id, frames, counters, checksum) -> np.ndarray:
    vals = [stream_id, frames, counters["n_seeds"], counters["n_solved"], counters["n_culled"], counters["n_fusions"],
            counters["bm_evals"], counters["lm_evals"], counters["map_size"], checksum]
    return np.array(vals, np.float64)


def reduce_and_gather(local_ms: float, local_evals: float, record: np.ndarray, device="cpu"):
    """Returns (max_ms, sum_evals, records[world, len(RECORD_FIELDS)]) on every rank.
    Works without an initialised process group (world size 1)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(local_ms), float(local_evals), record[None, :].copy()
    t = torch.tensor([local_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e = torch.tensor([local_evals], dtype=torch.float64, device=device)
    dist.all_reduce(e, op=dist.ReduceOp.SUM)
    rec = torch.from_numpy(record.copy()).to(device)
    out = [torch.empty_like(rec) for _ in range(dist.get_world_size())]
    dist.all_gather(out, rec)
    return float(t.item()), float(e.item()), np.stack([o.cpu().numpy() for o in out])


def gather_scalars(values, device="cpu"):
    """All-gather a short list of floats; returns an array [world, len(values)] on every rank (diagnostics)."""
    import torch
    import torch.distributed as dist
    v = np.asarray(values, np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return v[None, :].copy()
    t = torch.from_numpy(v.copy()).to(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return np.stack([o.cpu().numpy() for o in out])
