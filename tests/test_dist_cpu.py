"""world_size-2 `gloo` test of the sharding contract and the single metrics collective (CPU only).

The product env needs a GPU, so the per-rank env here is the oracle (test infrastructure) created with the same
(global_offset, global_n) arguments the product receives from b200bpp.dist.make_sharded_env."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, total, steps, out):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "online-3d-bpp-drl_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as orc
    from b200bpp import dist as bdist
    off, n = bdist.shard(total, rank, world)
    env = orc.OracleEnv(n, (10, 10, 10), rot=True, seed=21, global_offset=off, global_n=total, n_threads=1)
    env.reset()
    for _ in range(steps):
        env.step(env.sample.copy())
    red = bdist.reduce_metrics(env.metrics())
    plain, box, _, _ = env.get_state()
    if rank == 0:
        out.put(("metrics", red))
    out.put(("state", rank, off, plain, box))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_env():
    from b200bpp import dist as bdist
    import oracle as orc
    total, steps, world = 301, 20, 2          # odd total: uneven shards
    assert bdist.shard(total, 0, 2) == (0, 151) and bdist.shard(total, 1, 2) == (151, 150)
    assert sum(bdist.shard(10, r, 4)[1] for r in range(4)) == 10
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, steps, out)) for r in range(world)]
    [p.start() for p in procs]
    got = [out.get(timeout=120) for _ in range(world + 1)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    full = orc.OracleEnv(total, (10, 10, 10), rot=True, seed=21, n_threads=1)
    full.reset()
    for _ in range(steps):
        full.step(full.sample.copy())
    metrics = [g[1] for g in got if g[0] == "metrics"][0]
    assert metrics == full.metrics()
    fplain, fbox, _, _ = full.get_state()
    for g in got:
        if g[0] == "state":
            _, rank, off, plain, box = g
            assert np.array_equal(plain, fplain[off:off + len(plain)]) and np.array_equal(box, fbox[off:off + len(box)])


def test_episode_summary_and_single_process_reduce():
    from b200bpp import dist as bdist
    m = {"episodes": 4, "sum_counter": 24, "sum_vol": 1000, "steps": 100}
    assert bdist.reduce_metrics(m) == m                       # no process group -> identity
    s = bdist.episode_summary(m, (10, 10, 10))
    assert s["mean_items"] == 6 and abs(s["mean_ratio"] - 0.25) < 1e-12 and abs(s["mean_return"] - 2.5) < 1e-12
