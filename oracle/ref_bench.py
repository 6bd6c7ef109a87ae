"""Times the UNMODIFIED reference's own CPU path on this machine's host cores (BASELINE INFRASTRUCTURE).

What is timed is what the reference does per rollout step, with its own code (from /root/reference, or from the
byte-for-byte copy oracle/make_ref.py put under oracle/_ref/):

  shmem   `acktr.envs.make_vec_envs(...)` -> `ShmemVecEnv(context='fork')` of P `PackingGame` workers behind
          `VecNormalize` / `VecPyTorch` (acktr/envs.py:77-118), stepped with a uniformly random VALID action per env, plus
          the parent's row-by-row ground-truth mask loop (`get_possible_position` / `get_rotation_mask`, main.py:163-169).
          This is main.py's env side of a training step; it is the headline CPU figure.
  subproc `SubprocVecEnv(context='fork', in_series=k)` (baselines/common/vec_env/subproc_vec_env.py:39-138) over the same
          env thunks, step only and step + the same host mask loop -- the wording of north_star.

Nothing here is a product path.  bench.py's `--impl reference` arm and its `cpu_baseline` leg are the only callers.
"""
import contextlib
import io
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import reference_loader as rl   # noqa: E402


def host_cores():
    """(cores usable by this process, how that was determined): affinity mask capped by the cgroup CPU quota."""
    try:
        aff = len(os.sched_getaffinity(0))
    except AttributeError:
        aff = os.cpu_count() or 1
    quota = None
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    quota = float(txt[0]) / float(txt[1])
            else:
                q = float(txt[0])
                if q > 0:
                    quota = q / float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            break
        except (OSError, ValueError, IndexError):
            continue
    cores = aff if quota is None else max(1, min(aff, int(quota + 0.5)))
    return cores, {"affinity": aff, "cgroup_quota_cpus": quota}


class Args(object):
    """The fields of acktr/arguments.py:100-128 that the env side reads."""

    def __init__(self, container_size, enable_rotation, data_type, item_range=(2, 2, 2, 5, 5, 5)):
        self.container_size = tuple(container_size)
        self.enable_rotation = bool(enable_rotation)
        self.data_type = data_type
        r = item_range
        self.box_size_set = [(i, j, k) for i in range(r[0], r[3] + 1) for j in range(r[1], r[4] + 1)
                             for k in range(r[2], r[5] + 1)]


def _register():
    """main.py:224-227."""
    from gym.envs.registration import register
    with contextlib.suppress(Exception):
        register(id="Bpp-v0", entry_point="envs.bpp0:PackingGame")


def _mask_rows(utils, obs, args):
    """main.py:163-169, literally: one Python call per env row in the parent process."""
    import torch
    location_masks = []
    for observation in obs:
        if not args.enable_rotation:
            box_mask = utils.get_possible_position(observation, args.container_size)
        else:
            box_mask = utils.get_rotation_mask(observation, args.container_size)
        location_masks.append(box_mask)
    return torch.FloatTensor(np.asarray(location_masks))


def _valid_actions(masks, rng):
    """Uniformly random valid action per row (the policy stand-in of every arm of bench.py)."""
    import torch
    m = masks.numpy()
    u = rng.random(m.shape[0])
    cs = np.cumsum(m, 1)
    idx = (cs > (u * cs[:, -1])[:, None]).argmax(1)
    return torch.from_numpy(idx.astype(np.int64)).view(-1, 1)


def time_shmem(container_size=(10, 10, 10), rot=False, data_type="rs", processes=None, seconds=12.0, seed=1,
               steps=None, warmup_steps=None, with_mask=True):
    """-> dict(value = env-steps/s, ...).  `steps`: fixed number of bench steps, each `chunk` VecEnv.step calls."""
    ref = rl.load_reference()
    if ref is None:
        raise RuntimeError("neither /root/reference nor oracle/_ref is present: run oracle/make_ref.py in the build container")
    import torch
    torch.set_num_threads(1)                       # main.py:60
    cores, how = host_cores()
    P = int(processes or max(2, cores))
    args = Args(container_size, rot, data_type)
    _register()
    with rl.quiet():
        from acktr.envs import make_vec_envs
        log_dir = tempfile.mkdtemp(prefix="bpp_ref_log_")      # main.py:54-61: Monitor files per worker
        envs = make_vec_envs("Bpp-v0", seed, P, 0.99, log_dir, torch.device("cpu"), False, args=args)
        obs = envs.reset()
    rng = np.random.default_rng(seed)
    masks = _mask_rows(ref.utils, obs, args)
    episodes = 0

    def vec_step():
        nonlocal obs, masks, episodes
        action = _valid_actions(masks, rng)
        obs, reward, done, infos = envs.step(action)
        for i in range(len(infos)):                            # main.py:159-162
            if "episode" in infos[i].keys():
                episodes += 1
        if with_mask:
            masks = _mask_rows(ref.utils, obs, args)
        else:
            masks = torch.ones_like(masks)

    # calibrate: VecEnv.step calls per bench step so that one bench step is ~ seconds / steps
    t0 = time.perf_counter()
    n_cal = 0
    while time.perf_counter() - t0 < 1.0 or n_cal < 3:
        vec_step()
        n_cal += 1
    per_call = (time.perf_counter() - t0) / n_cal
    if steps is None:
        chunk, steps = max(1, int(0.5 / per_call)), None
    else:
        chunk = max(1, int(seconds / max(steps, 1) / per_call))
    for _ in range(int(warmup_steps or 0)):
        for _ in range(chunk):
            vec_step()
    t0 = time.perf_counter()
    done_steps = 0
    while True:
        for _ in range(chunk):
            vec_step()
        done_steps += 1
        if steps is not None:
            if done_steps >= steps:
                break
        elif time.perf_counter() - t0 >= seconds:
            break
    dt = time.perf_counter() - t0
    with rl.quiet():
        envs.close()
    return {"value": P * chunk * done_steps / dt, "unit": "env-steps/s", "cores": cores, "cores_detail": how, "kind": "reference",
            "processes": P, "bench_steps": done_steps, "vec_steps_per_bench_step": chunk, "seconds": dt, "episodes": episodes,
            "sample": "%d envs (ShmemVecEnv, fork) x %d VecEnv.step calls%s of the unmodified reference (%s, %s%s), %.1f s"
                      % (P, chunk * done_steps, " + host mask loop (main.py:163-169)" if with_mask else "", data_type,
                         "x".join(map(str, container_size)), ", rotation" if rot else "", dt)}


def time_subproc(container_size=(10, 10, 10), rot=False, data_type="rs", processes=None, in_series=4, seconds=6.0, seed=1,
                 with_mask=False):
    """SubprocVecEnv(context='fork', in_series=k) over the reference's own env thunks (acktr/envs.py:31-75)."""
    ref = rl.load_reference()
    if ref is None:
        raise RuntimeError("reference not present")
    import torch
    torch.set_num_threads(1)
    cores, how = host_cores()
    P = int(processes or max(2, cores))
    args = Args(container_size, rot, data_type)
    _register()
    with rl.quiet():
        from acktr.envs import make_env
        from baselines.common.vec_env.subproc_vec_env import SubprocVecEnv
        thunks = [make_env("Bpp-v0", seed, i, None, False, args) for i in range(P * in_series)]
        envs = SubprocVecEnv(thunks, context="fork", in_series=in_series)
        obs = envs.reset()
    n_env = P * in_series
    rng = np.random.default_rng(seed)
    masks = _mask_rows(ref.utils, torch.from_numpy(np.asarray(obs, np.float32)), args)
    t0 = time.perf_counter()
    calls = 0
    while time.perf_counter() - t0 < seconds or calls < 3:
        action = _valid_actions(masks, rng).view(-1).numpy()
        obs, reward, done, infos = envs.step(action)
        if with_mask:
            masks = _mask_rows(ref.utils, torch.from_numpy(np.asarray(obs, np.float32)), args)
        else:
            masks = torch.ones_like(masks)
        calls += 1
    dt = time.perf_counter() - t0
    with rl.quiet():
        envs.close()
    return {"value": n_env * calls / dt, "unit": "env-steps/s", "cores": cores, "processes": P, "in_series": in_series,
            "envs": n_env, "with_mask": with_mask, "seconds": dt}


if __name__ == "__main__":
    import json
    print(json.dumps(time_shmem(seconds=5.0)))
    print(json.dumps(time_subproc(seconds=3.0)))
    print(json.dumps(time_subproc(seconds=3.0, with_mask=True)))
