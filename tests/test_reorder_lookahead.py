"""Batched BPP-k lookahead (b200bpp/lookahead.py) against the reference's own search (`acktr/reorder.py`, unmodified).

Live test (needs the reference tree): four sequences of `dataset/cut_2.pt` are packed side by side; at every step the
reference runs `ReorderTree(...).reorder_search()` once per sequence on its own env, the batched search decides for all
of them in lockstep, and the two must agree on action, value and the "default" flag.  The placement engine of the batched
search is an oracle-backed stand-in here (no GPU in this container); the GPU test runs the same search on the CUDA env and
checks it against decisions recorded by this file's reference run (tests/golden/lookahead_cut2.npz).
"""
import os
import types

import numpy as np
import pytest
import torch

import oracle as orc
import reference_loader as rl
from b200bpp import lookahead
from b200bpp.policy import PackingPolicy

SIZE = (10, 10, 10)
PREVIEW = 4
TRAJS = (1, 2, 3, 5)


class OracleEngine(object):
    """The engine surface of BinPackingVecEnv that the lookahead uses, stepped by the CPU oracle (tests only)."""

    def __init__(self, n, size=SIZE):
        dummy = np.ones((1, 1, 3), dtype=np.uint8)
        self.o = orc.OracleEnv(n, size, rot=False, mask_variant=1, gen="replay", traj=dummy, traj_len=np.array([1], np.int32),
                               first_traj=0)
        self.o.reset()
        self.bin_size, self.device, self.ticks = tuple(size), torch.device("cpu"), 0

    def import_state(self, plain=None, box=None, counter=None, vol=None):
        self.o.set_state(plain=plain, box=box, counter=counter, vol=vol)

    def step_device(self, actions):
        self.ticks += 1
        self.o.step(actions.numpy().reshape(-1))
        return (torch.from_numpy(self.o.obs.copy()), torch.from_numpy(self.o.reward.copy()),
                torch.from_numpy(self.o.done.copy()), torch.from_numpy(self.o.mask.copy()))

    def export_state(self):
        return tuple(torch.from_numpy(a) for a in self.o.get_state())


def _state_of(env):
    """(plain, counter, vol) of a reference PackingGame."""
    return (np.asarray(env.space.plain, dtype=np.int32).reshape(-1), len(env.space.boxes),
            int(sum(b.x * b.y * b.z for b in env.space.boxes)))


@pytest.fixture(scope="module")
def weights(golden_dir):
    w = np.load(os.path.join(golden_dir, "policy_default_cut_2.npz"))
    return {k: w[k] for k in w.files}


def _reference_nmodel(weights):
    from acktr.model import Policy
    from acktr.model_loader import nnModel
    import gym
    args = types.SimpleNamespace(container_size=SIZE, enable_rotation=False, hidden_size=256, channel=4, pallet_size=10)
    pol = Policy(obs_shape=(400,), action_space=gym.spaces.Discrete(100),
                 base_kwargs={"recurrent": False, "hidden_size": 256, "args": args})
    pol.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    pol.eval()
    nm = nnModel.__new__(nnModel)
    nm.alen, nm.olen, nm.height, nm.device, nm._model = 100, 400, 10, torch.device("cpu"), pol
    return nm


def test_batched_lookahead_equals_reference_search(weights, golden_dir):
    REF = rl.load_reference()
    if REF is None:
        pytest.skip("reference tree not present (GPU box)")
    from acktr.reorder import ReorderTree
    torch.set_num_threads(1)
    nm = _reference_nmodel(weights)
    policy = PackingPolicy(SIZE).load_reference_state(weights)
    data = os.path.join(REF.path, "dataset", "cut_2.pt")
    item_set = [(i, j, k) for i in range(2, 6) for j in range(2, 6) for k in range(2, 6)]
    envs = []
    for t in TRAJS:
        with rl.quiet():
            e = REF.PackingGame(container_size=SIZE, box_set=item_set, test=True, data_name=data, enable_rotation=False)
        e.box_creator.reset(index=t)                       # LoadBoxCreator.reset(index): play trajectory t
        e.space = REF.Space(*SIZE)
        e.box_creator.generate_box_size()
        e.box_creator.preview(500)
        envs.append(e)
    n = len(envs)
    engine = OracleEngine(n)
    alive = [True] * n
    record = {"plain": [], "counter": [], "vol": [], "boxes": [], "act": [], "val": [], "default": [], "alive": []}
    steps = 0
    while any(alive) and steps < 40:
        states = [_state_of(e) for e in envs]
        box_lists = [e.box_creator.preview(PREVIEW) for e in envs]
        acts, vals, defaults = lookahead.reorder_search_batch(policy, engine, [s[0] for s in states], [s[1] for s in states],
                                                              [s[2] for s in states], box_lists, times=100)
        ref_act, ref_val, ref_def = [], [], []
        for i, e in enumerate(envs):
            if not alive[i]:                                # finished sequences stay in the batch, nobody reads them
                ref_act.append(0); ref_val.append(0.0); ref_def.append(True)
                continue
            with rl.quiet():
                a, v, d = ReorderTree(nm, box_lists[i], e, times=100).reorder_search()
            ref_act.append(int(a)); ref_val.append(float(v)); ref_def.append(bool(d))
            assert int(acts[i]) == int(a), (steps, i)
            assert bool(defaults[i]) == bool(d), (steps, i)
            assert vals[i] == pytest.approx(v, rel=1e-5, abs=1e-5), (steps, i)
        record["plain"].append(np.stack([s[0] for s in states])); record["counter"].append([s[1] for s in states])
        record["vol"].append([s[2] for s in states]); record["boxes"].append(np.asarray(box_lists, dtype=np.int32))
        record["act"].append(ref_act); record["val"].append(ref_val); record["default"].append(ref_def)
        record["alive"].append(list(alive))
        for i, e in enumerate(envs):
            if alive[i]:
                _, _, done, _ = e.step([ref_act[i]])
                alive[i] = not done
        steps += 1
    assert not any(alive) and steps > 10
    # a full search per sequence costs k + (k-1)! * k (evaluation, step) pairs at most; lockstep means the batch pays that
    # once per decision, not once per sequence
    assert engine.ticks <= steps * (PREVIEW + 6 * PREVIEW)
    if os.environ.get("BPP_WRITE_GOLDEN") == "1":          # regenerate the fixture of the GPU test below
        np.savez_compressed(os.path.join(golden_dir, "lookahead_cut2.npz"), preview=PREVIEW,
                            **{k: np.asarray(v) for k, v in record.items()})


@pytest.mark.gpu
def test_batched_lookahead_on_device(weights, golden_dir):
    """The same search with the CUDA env as placement engine and the network on the GPU, against the decisions the
    reference's ReorderTree took on these states (recorded by the test above with BPP_WRITE_GOLDEN=1)."""
    import b200bpp
    g = np.load(os.path.join(golden_dir, "lookahead_cut2.npz"))
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        dev = torch.device("cuda", 0)
        policy = PackingPolicy(SIZE).load_reference_state(weights).to(dev)
        n = g["act"].shape[1]
        dummy = (np.ones((1, 1, 3), dtype=np.uint8), np.array([1], dtype=np.int32))
        engine = b200bpp.BinPackingVecEnv(n, SIZE, trajectories=dummy, first_traj=0, mask_variant=1, device=dev, host_infos=False)
        engine.reset()
        agree = total = 0
        for t in range(g["act"].shape[0]):
            acts, vals, defaults = lookahead.reorder_search_batch(policy, engine, g["plain"][t], g["counter"][t], g["vol"][t],
                                                                  [list(map(tuple, b)) for b in g["boxes"][t]], times=100)
            for i in range(n):
                if g["alive"][t][i]:
                    total += 1
                    same = int(acts[i]) == int(g["act"][t][i]) and bool(defaults[i]) == bool(g["default"][t][i])
                    agree += same
                    if same:
                        assert abs(vals[i] - g["val"][t][i]) < 1e-3
        engine.close()
        # float32 convolutions reduce in another order on the GPU: a near-tie may flip one decision, not many
        assert total > 40 and agree >= total - 2, (agree, total)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
