"""Policy in the loop: the reference's actor-critic evaluated on the device tensors of BinPackingVecEnv.

SURVEY.md section 8(f) rank 3.  The network itself is NOT part of the hot path this repo rebuilds (it is plain
library convolutions / GEMMs through torch); it is restated here for inference only so that the pretrained weights
of the reference (`pretrained_models/*.pt`) can drive the new env batched on the GPU, which is the end-to-end check
that the env is behaviour-preserving: same observations + same masks => same greedy packing as the reference's
single-env loop (tests/golden/make_policy_golden.py records that loop from the unmodified reference).

What is restated, with the reference's arithmetic kept literally:
  * the `CNNPro` trunk and its three heads        acktr/model.py:262-323
  * the masked categorical head                   acktr/distributions.py:71-101
      probs = softmax(logits - (1 - mask) * 14) + 1e-5, mode = argmax, sample = Categorical(probs)
  * `Policy.act` / `Policy.get_value`             acktr/model.py:59-71,84-86
  * `nnModel.evaluate` (BPP-k scoring)            acktr/model_loader.py:39-68
      softmax(logits) * (pred_mask >= 0.5)

Training-side pieces (K-FAC wrappers `module.` / `add_bias.`, the three loss branches of Categorical.forward) are out
of scope; `load_reference_state` folds the K-FAC parameter naming back into plain weights and biases the same way
acktr/model_loader.py:27-33 does.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

MASK_LOGIT_PENALTY = 14.0          # acktr/distributions.py:87
PROB_FLOOR = 1e-5                  # acktr/distributions.py:88


class PackingPolicy(nn.Module):
    """Inference-only actor-critic over `float32[N, 4*L*L]` observations (channels: heightmap, x, y, z planes).

    The reference network assumes a square bin (`args.pallet_size = container_size[0]`, acktr/arguments.py:117).
    """

    # (attribute here, parameter prefix in the reference state dict)
    _REF_NAMES = (
        ("trunk.0", "base.share.0"), ("trunk.1", "base.share.2"), ("trunk.2", "base.share.4"),
        ("trunk.3", "base.share.6"), ("trunk.4", "base.share.8"),
        ("mask_conv", "base.mask.0"), ("mask_fc1", "base.mask.3"), ("mask_fc2", "base.mask.5"),
        ("actor_conv", "base.actor.0"), ("actor_fc", "base.actor.3"),
        ("critic_conv", "base.critic.0"), ("critic_fc", "base.critic.3"), ("critic_out", "base.critic_linear"),
        ("head", "dist.linear"),
    )

    def __init__(self, container_size=(10, 10, 10), enable_rotation=False, hidden_size=256, channel=4):
        super().__init__()
        L, W = int(container_size[0]), int(container_size[1])
        if L != W:
            raise ValueError("the reference network needs a square bin (pallet_size), got %dx%d" % (L, W))
        self.side, self.channel = L, int(channel)
        self.act_len = L * W * (2 if enable_rotation else 1)
        self.trunk = nn.ModuleList([nn.Conv2d(self.channel if i == 0 else 64, 64, 3, stride=1, padding=1)
                                    for i in range(5)])                        # model.py:267-278
        self.mask_conv = nn.Conv2d(64, 8, 1)                                    # model.py:283-292
        self.mask_fc1 = nn.Linear(8 * L * L, hidden_size)
        self.mask_fc2 = nn.Linear(hidden_size, self.act_len)
        self.actor_conv = nn.Conv2d(64, 8, 1)                                   # model.py:294-300
        self.actor_fc = nn.Linear(8 * L * L, hidden_size)
        self.critic_conv = nn.Conv2d(64, 4, 1)                                  # model.py:302-309
        self.critic_fc = nn.Linear(4 * L * L, hidden_size)
        self.critic_out = nn.Linear(hidden_size, 1)
        self.head = nn.Linear(hidden_size, self.act_len)                        # distributions.py:80
        self.eval()

    # -- weights -------------------------------------------------------------------------------------
    def load_reference_state(self, state):
        """`state`: the first element of a `pretrained_models/*.pt` pair, or the same tensors as numpy arrays
        (tests/golden/policy_default_cut_2.npz).  Keys may carry the K-FAC wrappers `module.` / `add_bias._bias`."""
        norm = {}
        for k, v in state.items():
            k = k.replace("module.", "").replace("add_bias.", "").replace("_bias", "bias")   # model_loader.py:27-29
            t = torch.as_tensor(v)
            if k.endswith("bias") and t.dim() == 2:                                          # (C, 1) -> (C), :31-33
                t = t.squeeze(-1)
            norm[k] = t
        own = {}
        for mine, ref in self._REF_NAMES:
            own[mine + ".weight"] = norm[ref + ".weight"]
            own[mine + ".bias"] = norm[ref + ".bias"]
        self.load_state_dict(own, strict=True)
        return self

    # -- network -------------------------------------------------------------------------------------
    def forward(self, obs):
        """-> (value [N,1], logits [N,M], pred_mask [N,M]); `CNNPro.forward` + the linear of the categorical head."""
        x = obs.reshape(-1, self.channel, self.side, self.side)                 # model.py:316
        for conv in self.trunk:
            x = F.relu(conv(x))
        n = x.shape[0]
        actor = F.relu(self.actor_fc(F.relu(self.actor_conv(x)).reshape(n, -1)))
        critic = F.relu(self.critic_fc(F.relu(self.critic_conv(x)).reshape(n, -1)))
        pm = F.relu(self.mask_fc1(F.relu(self.mask_conv(x)).reshape(n, -1)))
        pred_mask = F.relu(self.mask_fc2(pm))
        return self.critic_out(critic), self.head(actor), pred_mask

    @staticmethod
    def masked_probs(logits, location_masks):
        """acktr/distributions.py:82-92 (`fat_cat`): un-normalised; Categorical(probs=...) divides by the sum."""
        return F.softmax(logits - (1.0 - location_masks) * MASK_LOGIT_PENALTY, dim=-1) + PROB_FLOOR

    @torch.no_grad()
    def act(self, obs, location_masks, deterministic=False, generator=None):
        """`Policy.act` (model.py:59-71): -> (value [N,1], action int64 [N,1], log-prob [N,1])."""
        value, logits, _ = self.forward(obs)
        probs = self.masked_probs(logits, location_masks)
        if deterministic:
            action = probs.argmax(dim=-1, keepdim=True)                          # FixedCategorical.mode, :24
        else:
            action = torch.multinomial(probs, 1, generator=generator)
        logp = torch.log(probs / probs.sum(dim=-1, keepdim=True)).gather(1, action)
        return value, action, logp

    @torch.no_grad()
    def get_value(self, obs):
        return self.forward(obs)[0]

    @torch.no_grad()
    def evaluate(self, obs, use_mask=True):
        """`nnModel.evaluate` (model_loader.py:39-68), batched: -> (value [N], scores [N,M]) where
        scores = softmax(logits) * (pred_mask >= 0.5) -- the PREDICTED mask, not the ground truth."""
        value, logits, pred = self.forward(obs)
        scores = F.softmax(logits, dim=-1)
        if use_mask:
            scores = scores * (pred >= 0.5).to(scores.dtype)                     # Policy.binary, model.py:52-56
        return value.reshape(-1), scores


def greedy_policy(policy):
    """Adapter for rollout.collect / rollout.play_first_episodes: `(obs, masks) -> action [N,1]`."""
    def fn(obs, location_masks):
        return policy.act(obs, location_masks, deterministic=True)[1]
    return fn
