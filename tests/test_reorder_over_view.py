"""SURVEY.md 8(f) rank 4, first half: the reference's BPP-k lookahead (`acktr/reorder.py`, driven like
`unified_test.py:9-28`) runs UNMODIFIED on the single-env view `b200bpp.PackingGame` and takes the same decisions as on the
reference's own `envs.bpp0.PackingGame`.

The search deep-copies the env once per simulation, assigns `box_creator.box_list`, reads `cur_observation` and steps the
copies (reorder.py:69,164-166,190,250-254), so this exercises exactly the compatibility surface the view exists for.
Live test: needs the reference tree (build container only).  The view's device engine is replaced by an oracle-backed
stand-in here, because this container has no GPU; on a B200 the same view is covered against the golden traces by
tests/test_gpu_parity.py::test_single_env_view_matches_golden.
"""
import os
import types

import numpy as np
import pytest
import torch

import oracle as orc
import reference_loader as rl


class _OracleEngine(object):
    """What PackingGame needs from its 1-bin BinPackingVecEnv (state in, one step / one mask, state out), on the CPU."""

    def __init__(self, size, rot):
        dummy = np.ones((1, 1, 3), dtype=np.uint8)
        self.o = orc.OracleEnv(1, size, rot=rot, mask_variant=1, gen="replay", traj=dummy, traj_len=np.array([1], np.int32),
                               first_traj=0)
        self.o.reset()
        self.device = torch.device("cpu")

    def import_state(self, plain=None, box=None, counter=None, vol=None):
        self.o.set_state(plain=plain, box=box, counter=counter, vol=vol)

    def observe(self):
        obs, mask = self.o.observe()
        return torch.from_numpy(obs.copy()), torch.from_numpy(mask.copy())

    def step_device(self, actions):
        self.o.step(actions.numpy().reshape(-1))
        return (torch.from_numpy(self.o.obs.copy()), torch.from_numpy(self.o.reward.copy()),
                torch.from_numpy(self.o.done.copy()), torch.from_numpy(self.o.mask.copy()))

    def export_state(self):
        return tuple(torch.from_numpy(a) for a in self.o.get_state())

    def close(self):
        self.o.close()


@pytest.fixture()
def ref():
    REF = rl.load_reference()
    if REF is None:
        pytest.skip("reference tree not present (GPU box)")
    return REF


def _reference_nmodel(golden_dir):
    """acktr.model_loader.nnModel around the reference's own Policy with the pretrained tensors (fixture copy)."""
    from acktr.model import Policy
    from acktr.model_loader import nnModel
    import gym
    args = types.SimpleNamespace(container_size=(10, 10, 10), enable_rotation=False, hidden_size=256, channel=4, pallet_size=10)
    pol = Policy(obs_shape=(400,), action_space=gym.spaces.Discrete(100),
                 base_kwargs={"recurrent": False, "hidden_size": 256, "args": args})
    w = np.load(os.path.join(golden_dir, "policy_default_cut_2.npz"))
    pol.load_state_dict({k: torch.from_numpy(w[k]) for k in w.files})
    pol.eval()
    nm = nnModel.__new__(nnModel)                      # skip __init__: it torch.load()s a CUDA-saved checkpoint
    nm.alen, nm.olen, nm.height, nm.device, nm._model = 100, 400, 10, torch.device("cpu"), pol
    return nm


def test_reference_lookahead_runs_on_the_view(ref, golden_dir, monkeypatch):
    import copy
    from acktr.reorder import ReorderTree          # the reference's search, unmodified
    from b200bpp import bpp0
    torch.set_num_threads(1)
    size = (10, 10, 10)

    def oracle_engine(self):                         # stands in for the view's 1-bin BinPackingVecEnv
        if self._engine is None:
            self._engine = _OracleEngine(self.bin_size, self.can_rotate)
        return self._engine
    monkeypatch.setattr(bpp0.PackingGame, "_eng", oracle_engine)
    data = os.path.join(ref.path, "dataset", "cut_2.pt")
    item_set = [(i, j, k) for i in range(2, 6) for j in range(2, 6) for k in range(2, 6)]
    with rl.quiet():
        theirs = ref.PackingGame(container_size=size, box_set=item_set, test=True, data_name=data, enable_rotation=False)
    ours = bpp0.PackingGame(container_size=size, box_set=item_set, test=True, data_name=data, enable_rotation=False)
    nm = _reference_nmodel(golden_dir)
    preview, decisions = 4, 0
    for case in range(1):                            # unified_test.py:51-53: reset, preview(500), run_sequence
        theirs.reset(); ours.reset()
        theirs.box_creator.preview(500); ours.box_creator.preview(500)
        a_env, b_env = copy.deepcopy(theirs), copy.deepcopy(ours)
        assert np.array_equal(a_env.cur_observation, b_env.cur_observation)
        for _ in range(60):
            la, lb = a_env.box_creator.preview(preview), b_env.box_creator.preview(preview)
            assert [tuple(x) for x in la] == [tuple(x) for x in lb]
            with rl.quiet():
                act_a, val_a, def_a = ReorderTree(nm, la, a_env, times=100).reorder_search()
                act_b, val_b, def_b = ReorderTree(nm, lb, b_env, times=100).reorder_search()
            assert int(act_a) == int(act_b) and bool(def_a) == bool(def_b)
            assert val_a == pytest.approx(val_b, rel=1e-6, abs=1e-6)
            oa, ra, da, ia = a_env.step([act_a])
            ob, rb, db, ib = b_env.step([act_b])
            decisions += 1
            assert np.array_equal(oa, ob) and ra == pytest.approx(rb, abs=1e-12) and da == db
            assert ia["counter"] == ib["counter"] and ia["ratio"] == pytest.approx(ib["ratio"], abs=1e-12)
            assert np.array_equal(a_env.space.plain, b_env.space.plain)
            if da:
                break
        else:
            pytest.fail("episode did not end")
    assert decisions > 10
