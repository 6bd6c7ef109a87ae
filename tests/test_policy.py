"""Policy in the loop: the reference's pretrained actor-critic drives the batched env.

Fixtures come from the unmodified reference (tests/golden/make_policy_golden.py): the pretrained tensors of
`pretrained_models/default_cut_2.pt`, the reference network's outputs on sampled states, and the result of the
reference's own greedy single-env loop on `dataset/cut_2.pt` trajectories 1..64.

CPU tests: the network restatement (b200bpp/policy.py) against those outputs, live against `acktr.model.Policy` when the
reference tree is present, and the batched rollout helper over an oracle-backed stand-in env.
GPU test: the same rollout over the CUDA env -- 64 bins replaying trajectories 1..64 must pack exactly like the reference.
"""
import os
import types

import numpy as np
import pytest
import torch

from b200bpp.policy import PackingPolicy, greedy_policy
from b200bpp import rollout as ro


@pytest.fixture(scope="module")
def gold(golden_dir):
    w = dict(np.load(os.path.join(golden_dir, "policy_default_cut_2.npz")))
    r = dict(np.load(os.path.join(golden_dir, "policy_rollouts.npz")))
    d = np.load(os.path.join(golden_dir, "datasets.npz"))
    return types.SimpleNamespace(weights=w, roll=r, traj=d["cut_2"], traj_len=d["cut_2_len"])


@pytest.fixture(scope="module")
def policy(gold):
    return PackingPolicy((10, 10, 10), enable_rotation=False).load_reference_state(gold.weights)


def _close(a, b):
    return float(np.abs(a - b).max()) <= 2e-5 * float(np.abs(b).max()) + 2e-4


def test_network_matches_reference_outputs(policy, gold):
    """value / logits / predicted mask / greedy action / BPP-k scores of the reference network on 279 visited states."""
    torch.set_num_threads(1)
    r = gold.roll
    obs, mask = torch.from_numpy(r["obs"]), torch.from_numpy(r["mask"])
    with torch.no_grad():
        value, logits, pred = policy(obs)
    # float32 sums in a different blocking (batch of 279 here, one state at a time there); logits reach +-1300
    assert _close(value.numpy().reshape(-1), r["value"])
    assert _close(logits.numpy(), r["logits"])
    assert _close(pred.numpy(), r["pred_mask"])
    _, action, logp = policy.act(obs, mask, deterministic=True)
    assert np.array_equal(action.numpy().reshape(-1), r["action"])
    assert torch.all(logp <= 0) and action.dtype == torch.int64 and tuple(action.shape) == (obs.shape[0], 1)
    v2, scores = policy.evaluate(obs)
    sure = np.abs(r["pred_mask"] - 0.5) > 1e-3          # away from the binarisation threshold of Policy.binary
    assert np.allclose(scores.numpy()[sure], r["scores"][sure], rtol=2e-2, atol=1e-6)   # exp() of logits known to 5e-3
    assert np.array_equal(scores.numpy().argmax(1), r["scores"].argmax(1))
    assert _close(v2.numpy(), r["value"])
    # greedy actions are always ground-truth feasible here: the penalty of 14 on masked logits dominates
    assert np.all(r["mask"][np.arange(len(r["action"])), r["action"]] == 1.0)


def test_sampling_respects_the_reference_distribution(policy, gold):
    """Stochastic `act`: samples follow softmax(logits - 14*(1-mask)) + 1e-5, normalised (distributions.py:82-92)."""
    r = gold.roll
    obs = torch.from_numpy(r["obs"][:1]).repeat(20000, 1)
    mask = torch.from_numpy(r["mask"][:1]).repeat(20000, 1)
    g = torch.Generator().manual_seed(3)
    _, a, logp = policy.act(obs, mask, deterministic=False, generator=g)
    with torch.no_grad():
        p = PackingPolicy.masked_probs(policy(obs[:1])[1], mask[:1])[0]
    p = (p / p.sum()).numpy()
    freq = np.bincount(a.numpy().reshape(-1), minlength=p.shape[0]) / a.shape[0]
    assert np.abs(freq - p).max() < 0.02
    assert np.allclose(np.exp(logp.numpy().reshape(-1)), p[a.numpy().reshape(-1)], rtol=1e-5)


def test_live_against_reference_policy(policy, gold):
    """Same tensors through acktr.model.Policy (unmodified reference), incl. a batch > 1 and the sampled branch's log-probs."""
    import reference_loader as rl
    if rl.load_reference() is None:
        pytest.skip("reference tree not present (GPU box)")
    from acktr.model import Policy
    import gym
    args = types.SimpleNamespace(container_size=(10, 10, 10), enable_rotation=False, hidden_size=256, channel=4,
                                 pallet_size=10)
    ref = Policy(obs_shape=(400,), action_space=gym.spaces.Discrete(100),
                 base_kwargs={"recurrent": False, "hidden_size": 256, "args": args})
    ref.load_state_dict({k: torch.from_numpy(v) for k, v in gold.weights.items()})
    ref.eval()
    obs, mask = torch.from_numpy(gold.roll["obs"][:64]), torch.from_numpy(gold.roll["mask"][:64])
    with torch.no_grad():
        v_ref, a_ref, lp_ref, _ = ref.act(obs, None, None, mask, deterministic=True)
    v, a, lp = policy.act(obs, mask, deterministic=True)
    assert torch.equal(a, a_ref)
    assert torch.allclose(v, v_ref, rtol=1e-5, atol=1e-5) and torch.allclose(lp, lp_ref, rtol=1e-5, atol=1e-5)
    assert torch.allclose(policy.get_value(obs), ref.get_value(obs, None, None), rtol=1e-5, atol=1e-5)


class _OracleBackedEnv(object):
    """Stand-in with BinPackingVecEnv's device-resident surface, stepped by the CPU oracle (tests only)."""

    def __init__(self, n, gold, first_traj=1):
        import oracle as orc
        self.o = orc.OracleEnv(n, (10, 10, 10), gen="replay", traj=gold.traj, traj_len=gold.traj_len,
                               first_traj=first_traj, n_threads=1)
        self.num_envs, self.device = n, torch.device("cpu")

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


def test_batched_greedy_rollout_equals_reference_loop_cpu(policy, gold):
    """64 bins in one batch (bin g replays trajectory 1+g) == the reference's one-sequence-at-a-time loop."""
    torch.set_num_threads(1)
    n = len(gold.roll["counter"])
    env = _OracleBackedEnv(n, gold, first_traj=int(gold.roll["first_traj"]))
    counter, ratio = ro.play_first_episodes(env, greedy_policy(policy), max_steps=200, check_every=4)
    assert np.array_equal(counter.numpy(), gold.roll["counter"])
    assert np.allclose(ratio.numpy(), gold.roll["ratio"], atol=1e-6)


@pytest.mark.gpu
def test_policy_in_the_loop_on_device(gold):
    """The CUDA env + the network on the GPU reproduce the reference's greedy packing of cut_2.pt trajectories 1..64,
    and a large batch (every trajectory of the dataset several times over) gives the same mean utilisation per trajectory."""
    import b200bpp
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False           # float32 convolutions / GEMMs, as on the reference's CPU path
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        dev = torch.device("cuda", 0)
        pol = PackingPolicy((10, 10, 10)).load_reference_state(gold.weights).to(dev)
        n = len(gold.roll["counter"])
        env = b200bpp.BinPackingVecEnv(n, (10, 10, 10), trajectories=(gold.traj, gold.traj_len),
                                       first_traj=int(gold.roll["first_traj"]), device=dev)
        counter, ratio = ro.play_first_episodes(env, greedy_policy(pol), max_steps=200)
        env.close()
        c, r = counter.cpu().numpy(), ratio.cpu().numpy()
        # float32 reductions run in another order on the GPU: a near-tie between two logits may flip an argmax
        same = c == gold.roll["counter"]
        assert same.mean() >= 0.95, "only %d of %d sequences packed like the reference" % (same.sum(), n)
        assert np.allclose(r[same], gold.roll["ratio"][same], atol=1e-6)
        assert abs(r.mean() - gold.roll["ratio"].mean()) < 0.01

        T = gold.traj.shape[0]
        big = b200bpp.BinPackingVecEnv(4 * T, (10, 10, 10), trajectories=(gold.traj, gold.traj_len), first_traj=1, device=dev)
        cb, rb = ro.play_first_episodes(big, greedy_policy(pol), max_steps=200)
        big.close()
        rb = rb.cpu().numpy().reshape(4, T)             # bin g plays trajectory (1 + g) mod T: four copies of each
        assert (rb[0] == rb[1]).mean() > 0.99 and (rb[0] == rb[3]).mean() > 0.99   # same sequence, same packing
        assert (rb[0][:n][same] == r[same]).mean() > 0.95
        assert 0.60 < rb.mean() < 0.75                  # the reference reports 66.9 % for this model without lookahead
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


def test_rotation_model_live(golden_dir):
    """`pretrained_models/rotation_cut_2.pt` (action space 2*L*W) through the restated network and through the reference's
    `Policy`, on states of a recorded rotation rollout: same greedy actions, values and log-probabilities."""
    import reference_loader as rl
    REF = rl.load_reference()
    if REF is None:
        pytest.skip("reference tree not present (GPU box)")
    from acktr.model import Policy
    import gym
    state, _ = torch.load(os.path.join(REF.path, "pretrained_models", "rotation_cut_2.pt"), map_location="cpu", weights_only=False)
    pol = PackingPolicy((10, 10, 10), enable_rotation=True).load_reference_state(state)
    args = types.SimpleNamespace(container_size=(10, 10, 10), enable_rotation=True, hidden_size=256, channel=4, pallet_size=10)
    ref = Policy(obs_shape=(400,), action_space=gym.spaces.Discrete(200),
                 base_kwargs={"recurrent": False, "hidden_size": 256, "args": args})
    folded = {k.replace("module.", "").replace("add_bias.", "").replace("_bias", "bias"): v for k, v in state.items()}
    ref.load_state_dict({k: (v.squeeze(-1) if v.dim() <= 3 else v) for k, v in folded.items()})      # model_loader.py:27-33
    ref.eval()
    g = np.load(os.path.join(golden_dir, "ref_rollouts_10_rot.npz"))
    hm, box = g["hm"][:8].reshape(-1, 100), g["box"][:8].reshape(-1, 3)
    obs = torch.from_numpy(np.concatenate([hm] + [np.repeat(box[:, c:c + 1], 100, 1) for c in range(3)], 1).astype(np.float32))
    mask = torch.from_numpy(g["mask_v2"][:8].reshape(-1, 200).astype(np.float32))
    with torch.no_grad():
        v_ref, a_ref, lp_ref, _ = ref.act(obs, None, None, mask, deterministic=True)
    v, a, lp = pol.act(obs, mask, deterministic=True)
    assert torch.equal(a, a_ref) and int((a >= 100).sum()) > 0              # rotated placements do get chosen
    assert torch.allclose(v, v_ref, rtol=1e-5, atol=1e-5) and torch.allclose(lp, lp_ref, rtol=1e-5, atol=1e-5)
