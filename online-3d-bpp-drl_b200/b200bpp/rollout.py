"""Device-resident rollout loop: the hot loop of main.py:148-174 without leaving the GPU.

The reference's loop calls `envs.step(action)` (two process hops, two PCIe hops), rebuilds the mask row by row on the
host (main.py:163-169) and builds `masks` / `bad_masks` with Python list comprehensions (main.py:172-173).  Here the
env hands back CUDA tensors, so the storage below (same tensors as acktr/storage.py:9-47: obs[T+1,N,4LW],
location_masks[T+1,N,LW(1+rot)], rewards[T,N,1], masks[T+1,N,1], actions[T,N,1]) is filled with device copies only.
Episode statistics (what baselines/bench/monitor.py:58-75 puts into info['episode'] and main.py:159-162 collects)
are kept as device tensors and read back only when asked.
"""
import torch


class DeviceRolloutStorage(object):
    def __init__(self, num_steps, env):
        N, dev = env.num_envs, env.device
        self.num_steps = num_steps
        self.obs = torch.zeros(num_steps + 1, N, env.obs_len, device=dev)
        self.location_masks = torch.zeros(num_steps + 1, N, env.act_len, device=dev)
        self.rewards = torch.zeros(num_steps, N, 1, device=dev)
        self.masks = torch.ones(num_steps + 1, N, 1, device=dev)          # 0 where the episode just ended
        self.bad_masks = torch.ones(num_steps + 1, N, 1, device=dev)      # never set by this env (main.py:173)
        self.actions = torch.zeros(num_steps, N, 1, dtype=torch.int64, device=dev)
        self.step = 0

    def insert(self, obs, actions, rewards, done, location_masks):
        t = self.step
        self.obs[t + 1].copy_(obs)
        self.location_masks[t + 1].copy_(location_masks)
        self.actions[t].copy_(actions.view(-1, 1))
        self.rewards[t].copy_(rewards.view(-1, 1))
        self.masks[t + 1].copy_((1 - done.view(-1, 1)).to(self.masks.dtype))
        self.step = (t + 1) % self.num_steps

    def after_update(self):
        """acktr/storage.py:103-108"""
        self.obs[0].copy_(self.obs[-1])
        self.location_masks[0].copy_(self.location_masks[-1])
        self.masks[0].copy_(self.masks[-1])


class EpisodeStats(object):
    """Monitor-equivalent on the device: sums of finished-episode return / length / ratio since the last read."""

    def __init__(self, env):
        dev = env.device
        self.lwh = float(env.bin_size[0] * env.bin_size[1] * env.bin_size[2])
        self.acc = torch.zeros(4, dtype=torch.float64, device=dev)     # episodes, sum return, sum length, sum ratio

    def update(self, done, counter, ratio):
        d = done.to(torch.float64)
        # `ratio` is float32(vol / LWH): the placed volume is an integer below 2^23, so rounding ratio * LWH recovers it
        # exactly and the statistics are those of the reference's float64 arithmetic (space.py:146-151, monitor.py:64)
        r = torch.round(ratio.to(torch.float64) * self.lwh) / self.lwh
        # episode return = 10 * ratio (sum of 10*v_i/LWH); length counts the failing step (monitor.py:62-64)
        self.acc += torch.stack([d.sum(), (d * r * 10.0).sum(), (d * (counter.to(torch.float64) + 1)).sum(), (d * r).sum()])

    def read(self, reset=True):
        n, ret, length, ratio = self.acc.tolist()
        if reset:
            self.acc.zero_()
        n1 = max(n, 1.0)
        return {"episodes": int(n), "mean_return": ret / n1, "mean_length": length / n1, "mean_ratio": ratio / n1}


def uniform_valid_policy(obs, location_masks):
    """Stand-in for actor_critic.act: a uniformly random valid action per bin (mask rows always have a 1)."""
    return torch.multinomial(location_masks, 1)


def collect(env, storage, policy=uniform_valid_policy, stats=None):
    """One rollout segment of storage.num_steps steps (main.py:148-174).  `policy(obs, masks) -> int64 [N,1]`."""
    for t in range(storage.num_steps):
        with torch.no_grad():
            action = policy(storage.obs[t], storage.location_masks[t])
        obs, reward, done, masks = env.step_device(action)
        if stats is not None:
            stats.update(done, env.counter_t, env.ratio_t)
        storage.insert(obs, action, reward, done, masks)
    return storage


def start(env, storage):
    """main.py:121-133: first observation and mask into slot 0."""
    obs = env.reset()
    storage.obs[0].copy_(obs)
    storage.location_masks[0].copy_(env.location_masks)
    storage.step = 0
    return storage


def play_first_episodes(env, policy, max_steps=1000, check_every=8):
    """Resets `env` and plays every bin's FIRST episode to its end with `policy(obs, masks) -> action [N,1]`; the batched
    form of the reference's test loop (one `env.reset()` + `env.step([act])` until done per sequence, unified_test.py:9-28
    with --preview 1, evaluation.py:20-41).  Bins that finish keep stepping through their auto-reset episodes (their
    results are ignored) until all are done.  -> (counter int32 [N], ratio float32 [N]) of the first episodes, on the device.
    """
    N, dev = env.num_envs, env.device
    first_done = torch.zeros(N, dtype=torch.bool, device=dev)
    counter = torch.zeros(N, dtype=torch.int32, device=dev)
    ratio = torch.zeros(N, dtype=torch.float32, device=dev)
    obs = env.reset()
    masks = env.location_masks
    for t in range(max_steps):
        with torch.no_grad():
            action = policy(obs, masks)
        obs, _, done, masks = env.step_device(action)
        newly = done.bool() & ~first_done
        counter = torch.where(newly, env.counter_t, counter)
        ratio = torch.where(newly, env.ratio_t, ratio)
        first_done |= newly
        if (t + 1) % check_every == 0 and bool(first_done.all()):       # the only host sync of the loop
            break
    if not bool(first_done.all()):
        raise RuntimeError("play_first_episodes: %d bins still running after %d steps"
                           % (int((~first_done).sum()), max_steps))
    return counter, ratio
