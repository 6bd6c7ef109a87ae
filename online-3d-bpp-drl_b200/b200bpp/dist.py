"""Multi-GPU plumbing: bins shard embarrassingly, one process per GPU (torchrun), no data-path collective.

The reference's only parallelism is env data-parallelism over OS processes (acktr/envs.py:86-101); here rank r
owns the global bins [r*n, (r+1)*n).  Global bin ids seed the Philox streams and pick replay trajectories, so the
job's results do not depend on the number of ranks.  The single collective is the all-reduce of the four integer
episode accumulators (what main.py:159-162 gathers from infos), issued once per log interval.
"""
import os

import torch


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard(total_bins, rank=None, world=None):
    """Contiguous block partition -> (global_offset, n_local).  total_bins need not divide evenly."""
    if rank is None or world is None:
        rank, world = rank_world()
    base, extra = divmod(int(total_bins), int(world))
    n = base + (1 if rank < extra else 0)
    off = rank * base + min(rank, extra)
    return off, n


def make_sharded_env(total_bins, rank=None, world=None, device=None, **kwargs):
    """BinPackingVecEnv over this rank's shard of a `total_bins` job."""
    from .vec_env import BinPackingVecEnv
    if rank is None or world is None:
        rank, world = rank_world()
    off, n = shard(total_bins, rank, world)
    if device is None:
        device = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
    return BinPackingVecEnv(n, global_offset=off, global_n=total_bins, device=device, **kwargs)


def reduce_metrics(local, group=None, device=None):
    """Sum {'episodes','sum_counter','sum_vol','steps'} over ranks (one all-reduce of 4 int64)."""
    import torch.distributed as dist
    keys = ("episodes", "sum_counter", "sum_vol", "steps")
    if not (dist.is_available() and dist.is_initialized()):
        return dict(local)
    if device is None:
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([int(local[k]) for k in keys], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return {k: int(v) for k, v in zip(keys, t.tolist())}


def episode_summary(m, container_size):
    """Mean items per episode and mean space utilisation (info['counter'], info['ratio'] averaged)."""
    L, W, H = container_size
    e = max(int(m["episodes"]), 1)
    return {"episodes": int(m["episodes"]), "mean_items": m["sum_counter"] / e, "mean_ratio": m["sum_vol"] / e / (L * W * H),
            "mean_return": 10.0 * m["sum_vol"] / e / (L * W * H)}
