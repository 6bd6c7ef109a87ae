"""The device rollout glue (b200bpp/rollout.py: DeviceRolloutStorage, collect, EpisodeStats) against the reference's own
rollout loop (main.py:121-174) over the reference's own env stack -- `DummyVecEnv` of `PackingGame`s behind `Monitor`,
`VecNormalize` (pass-through) and `VecPyTorch`, masks rebuilt row by row on the host, `acktr.storage.RolloutStorage`.

Live test (needs the reference tree).  Our side steps an oracle-backed stand-in with BinPackingVecEnv's device-resident
surface, because this container has no GPU; the GPU twin is tests/test_gpu_parity.py::test_device_rollout_glue.
"""
import os

import numpy as np
import pytest
import torch

import oracle as orc
import reference_loader as rl
from b200bpp import rollout

SIZE = (10, 10, 10)
N, T_SEG, SEGMENTS = 6, 5, 7


class OracleVecEnv(object):
    """BinPackingVecEnv's device-resident surface on the CPU (tests only)."""

    def __init__(self, n, traj, traj_len, rot):
        self.o = orc.OracleEnv(n, SIZE, rot=rot, gen="replay", traj=traj, traj_len=traj_len, first_traj=1)
        self.num_envs, self.device, self.bin_size = n, torch.device("cpu"), SIZE
        self.obs_len, self.act_len = 400, 100 * (2 if rot else 1)

    def _sync(self):
        self.obs = torch.from_numpy(self.o.obs.copy())
        self.location_masks = torch.from_numpy(self.o.mask.copy())
        self.counter_t = torch.from_numpy(self.o.counter.copy())
        self.ratio_t = torch.from_numpy(self.o.ratio.copy())

    def reset(self):
        self.o.reset()
        self._sync()
        return self.obs

    def step_device(self, actions):
        self.o.step(actions.numpy().reshape(-1))
        self._sync()
        return self.obs, torch.from_numpy(self.o.reward.copy()), torch.from_numpy(self.o.done.copy()), self.location_masks


@pytest.mark.parametrize("rot", [False, True])
def test_collect_equals_reference_rollout_loop(golden_dir, rot):
    REF = rl.load_reference()
    if REF is None:
        pytest.skip("reference tree not present (GPU box)")
    import gym
    from acktr.envs import VecNormalize, VecPyTorch
    from acktr.storage import RolloutStorage
    from baselines import bench
    from baselines.common.vec_env.dummy_vec_env import DummyVecEnv
    d = np.load(os.path.join(golden_dir, "datasets.npz"))
    traj, traj_len = d["cut_2"], d["cut_2_len"]
    Tn = traj.shape[0]

    class Scripted(REF.BoxCreator):
        """LoadBoxCreator semantics (binCreator.py:51-72) for bin g of N: episode e plays trajectory (1 + g + e*N) mod T."""

        def __init__(self, g):
            super().__init__()
            self.g, self.e, self.k = g, -1, 0

        def reset(self):
            self.box_list.clear()
            self.e += 1
            row = (1 + self.g + self.e * N) % Tn
            self.items = [tuple(int(v) for v in traj[row, k]) for k in range(int(traj_len[row]))]
            self.k = 0

        def generate_box_size(self, **kw):
            self.box_list.append(self.items[self.k] if self.k < len(self.items) else (10, 10, 10))
            self.k += 1

    def thunk(g):
        def make():
            with rl.quiet():
                env = REF.PackingGame(box_creator=Scripted(g), container_size=SIZE, enable_rotation=rot)
            return bench.Monitor(env, None, allow_early_resets=False)           # acktr/envs.py:55-59
        return make

    with rl.quiet():
        venv = VecPyTorch(VecNormalize(DummyVecEnv([thunk(g) for g in range(N)]), gamma=1.0, ob=False, ret=False), "cpu")
    act_len = 100 * (2 if rot else 1)
    ref_store = RolloutStorage(T_SEG, N, (400,), gym.spaces.Discrete(act_len), 1, can_give_up=False, enable_rotation=rot,
                               pallet_size=10)

    def host_masks(obs):                                                        # main.py:123-129,163-169
        rows = [REF.utils.get_rotation_mask(o, SIZE) if rot else REF.utils.get_possible_position(o, SIZE) for o in obs]
        return torch.FloatTensor(np.asarray(rows))

    # a deterministic stand-in for actor_critic.act: the valid action with the largest fixed pseudo-random weight
    weights = torch.rand(SEGMENTS * T_SEG, N, act_len, generator=torch.Generator().manual_seed(7))
    clock = {"ours": 0, "ref": 0}

    def policy_for(who):
        def policy(obs, masks):
            w = weights[clock[who]]
            clock[who] += 1
            return (masks * (1.0 + w)).argmax(dim=1, keepdim=True)
        return policy

    ours_env = OracleVecEnv(N, traj, traj_len, rot)
    ours_store = rollout.start(ours_env, rollout.DeviceRolloutStorage(T_SEG, ours_env))
    stats = rollout.EpisodeStats(ours_env)

    obs = venv.reset()
    masks = host_masks(obs)
    ref_store.obs[0].copy_(obs)
    ref_store.location_masks[0].copy_(masks)
    ref_policy = policy_for("ref")
    ep_r, ep_l, ep_ratio = [], [], []
    for seg in range(SEGMENTS):
        for step in range(T_SEG):                                               # main.py:148-174
            action = ref_policy(ref_store.obs[step], masks)
            obs, reward, done, infos = venv.step(action)
            for info in infos:
                if "episode" in info:
                    ep_r.append(info["episode"]["r"]); ep_l.append(info["episode"]["l"]); ep_ratio.append(info["ratio"])
            masks = host_masks(obs)
            m = torch.FloatTensor([[0.0] if d_ else [1.0] for d_ in done])
            bad = torch.FloatTensor([[0.0] if "bad_transition" in info.keys() else [1.0] for info in infos])
            ref_store.insert(obs, torch.zeros(N, 1), action, torch.zeros(N, 1), torch.zeros(N, 1), reward, m, bad, masks)
        rollout.collect(ours_env, ours_store, policy=policy_for("ours"), stats=stats)
        for name in ("obs", "location_masks", "rewards", "masks", "bad_masks", "actions"):
            assert torch.equal(getattr(ours_store, name), getattr(ref_store, name)), (seg, name)
        ref_store.after_update()
        ours_store.after_update()
        assert torch.equal(ours_store.obs[0], ref_store.obs[0]) and torch.equal(ours_store.masks[0], ref_store.masks[0])
    s = stats.read()
    assert s["episodes"] == len(ep_r) and len(ep_r) >= 2
    assert s["mean_return"] == pytest.approx(np.mean(ep_r), abs=1e-5)           # Monitor's r = sum of rewards (rounded to 1e-6)
    assert s["mean_length"] == pytest.approx(np.mean(ep_l), abs=1e-9)           # l counts the failing step
    assert s["mean_ratio"] == pytest.approx(np.mean(ep_ratio), abs=1e-6)
