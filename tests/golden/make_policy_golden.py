#!/usr/bin/env python
"""Records the reference's policy-in-the-loop behaviour with the UNMODIFIED Python reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_policy_golden.py

For `pretrained_models/default_cut_2.pt` (10x10x10, no rotation) it runs the reference's own pieces end to end --
`envs.bpp0.PackingGame(test=True, data_name=dataset/cut_2.pt)` (LoadBoxCreator: reset k plays trajectory k),
`acktr.utils.get_possible_position` for the ground-truth mask (main.py:163-169) and
`acktr.model.Policy.act(..., deterministic=True)` (evaluation.py:27-31) -- one sequence at a time, as
unified_test.py / evaluation.py do, and writes

  policy_default_cut_2.npz   the pretrained tensors (names with the K-FAC wrappers folded, model_loader.py:27-33);
                             the GPU box has no reference tree, so the weights travel as a fixture
  policy_rollouts.npz        per trajectory 1..K: items placed (info['counter']) and utilisation (info['ratio']) of
                             the greedy rollout; plus a sample of (obs, mask) pairs with the reference network's
                             value / logits / predicted mask / greedy action / nnModel.evaluate scores on them.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import reference_loader as rl   # noqa: E402

REF = rl.load_reference()
if REF is None:
    sys.exit("reference tree not found; fixtures can only be regenerated in the build container")

from acktr.model import Policy            # noqa: E402  (the reference's network, unmodified)
from acktr.model_loader import nnModel    # noqa: E402
import gym                                # noqa: E402  (oracle/ref_shim)

K_TRAJ = 64
SIZE = (10, 10, 10)


def reference_args():
    a = types.SimpleNamespace()
    a.container_size, a.enable_rotation, a.hidden_size = SIZE, False, 256
    a.channel, a.pallet_size, a.device = 4, SIZE[0], "cpu"          # acktr/arguments.py:117-118
    return a


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    args = reference_args()
    url = os.path.join(REF.path, "pretrained_models", "default_cut_2.pt")
    _orig_load = torch.load
    torch.load = lambda f, *a, **k: _orig_load(f, *a, **{**k, "weights_only": False, "map_location": "cpu"})   # torch >= 2.6 default; CUDA-saved tensors
    with rl.quiet():
        nmodel = nnModel(url, args)                                  # reference loader: key folding + Policy(...)
    model = nmodel._model.eval()
    assert isinstance(model, Policy)
    state = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    np.savez_compressed(os.path.join(HERE, "policy_default_cut_2.npz"), **state)

    item_set = [(i, j, k) for i in range(2, 6) for j in range(2, 6) for k in range(2, 6)]
    with rl.quiet():
        env = REF.PackingGame(container_size=SIZE, box_set=item_set, test=True, enable_rotation=False,
                              data_name=os.path.join(REF.path, "dataset", "cut_2.pt"))
    counters, ratios, lengths = [], [], []
    s_obs, s_mask, s_value, s_logits, s_pred, s_action, s_scores = [], [], [], [], [], [], []
    for k in range(1, K_TRAJ + 1):
        obs = env.reset()                                            # LoadBoxCreator.index += 1 -> trajectory k
        assert env.box_creator.index == k
        t = 0
        while True:
            mask = np.asarray(REF.utils.get_possible_position(obs.astype(np.float32), SIZE), dtype=np.float32)
            ot = torch.from_numpy(obs.astype(np.float32)).unsqueeze(0)
            mt = torch.from_numpy(mask).unsqueeze(0)
            with torch.no_grad():
                value, action, _, _ = model.act(ot, None, None, mt, deterministic=True)
                if k <= 16:                                          # numerics sample: every step of 16 sequences
                    v2, feat, _, pred = model.base(ot, None, None)
                    logits = model.dist.get_policy_distribution(feat)
                    _, scores = nmodel.evaluate(obs.astype(np.float32))
                    s_obs.append(obs.astype(np.float32)); s_mask.append(mask)
                    s_value.append(float(v2)); s_logits.append(logits[0].numpy()); s_pred.append(pred[0].numpy())
                    s_action.append(int(action)); s_scores.append(np.asarray(scores, dtype=np.float32))
            obs, reward, done, info = env.step([int(action)])
            t += 1
            if done:
                counters.append(info["counter"]); ratios.append(info["ratio"]); lengths.append(t)
                break
    np.savez_compressed(
        os.path.join(HERE, "policy_rollouts.npz"),
        first_traj=np.int32(1), counter=np.asarray(counters, np.int32), ratio=np.asarray(ratios, np.float64),
        steps=np.asarray(lengths, np.int32),
        obs=np.asarray(s_obs, np.float32), mask=np.asarray(s_mask, np.float32), value=np.asarray(s_value, np.float32),
        logits=np.asarray(s_logits, np.float32), pred_mask=np.asarray(s_pred, np.float32),
        action=np.asarray(s_action, np.int64), scores=np.asarray(s_scores, np.float32))
    print("reference greedy policy on cut_2.pt trajectories 1..%d: mean items %.3f, mean utilisation %.4f; %d sampled states"
          % (K_TRAJ, float(np.mean(counters)), float(np.mean(ratios)), len(s_obs)))


if __name__ == "__main__":
    main()
