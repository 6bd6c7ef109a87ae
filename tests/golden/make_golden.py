#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ by running the UNMODIFIED Python reference.

Run in the build container only (needs /root/reference; the GPU box has no reference tree):

    python tests/golden/make_golden.py

Outputs (all small, committed):
  ref_rollouts_<tag>.npz   VecEnv-style traces (auto-reset on done, baselines shmem_vec_env.py:127-130)
                           of envs.bpp0.PackingGame with injected item sequences and a shared action
                           stream: box, heightmap, mask v1 (Space.check_box), mask v2
                           (acktr.utils.get_possible_position / get_rotation_mask), reward, done, info.
  traj0.npz                user_study/trajs/traj0.txt (the repo's only recorded game).
  cut2_txt_sample.npz      first trajectories of user_study/cut_2.txt (pieces WITH positions).
  datasets.npz             dataset/cut_1.pt, cut_2.pt, rs.pt re-encoded as padded uint8 arrays.
  gen_stats.npz            statistics of the reference's cut1 / cut2 / rs generators (MT19937-driven),
                           used for distributional parity of the Philox-driven restatements.
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import reference_loader as rl   # noqa: E402

REF = rl.load_reference()
if REF is None:
    sys.exit("reference tree not found; fixtures can only be regenerated in the build container")


class ScriptedBoxCreator(REF.BoxCreator):
    """Injection hook (PackingGame(box_creator=...), bin3D.py:10-13) with LoadBoxCreator semantics
    (binCreator.py:51-72): episode e of this bin plays trajs[e]; past its end (10,10,10) forever."""

    def __init__(self, trajs):
        super().__init__()
        self.trajs = trajs
        self.index = -1
        self.box_index = 0

    def reset(self, index=None):
        self.box_list.clear()
        self.index += 1
        self.box_set = [tuple(int(v) for v in b) for b in self.trajs[self.index]]
        self.box_index = 0

    def generate_box_size(self, **kwargs):
        if self.box_index < len(self.box_set):
            self.box_list.append(self.box_set[self.box_index])
        else:
            self.box_list.append((10, 10, 10))
        self.box_index += 1


def mask_v1(env, rot):
    """Env-side mask (bin3D.py:72-93) and its rotation analogue built from Space.check_box."""
    L, W = int(env.space.plain_size[0]), int(env.space.plain_size[1])
    bx, by, bz = env.next_box
    halves = []
    for (x, y) in ([(bx, by), (by, bx)] if rot else [(bx, by)]):
        m = np.zeros((L, W), dtype=np.int32)
        for i in range(L - x + 1):
            for j in range(W - y + 1):
                if env.space.check_box(env.space.plain, x, y, i, j, bz) >= 0:
                    m[i, j] = 1
        halves.append(m.reshape(-1))
    m = np.hstack(halves)
    if m.sum() == 0:
        m[:] = 1
    if not rot:
        assert np.array_equal(m, env.get_possible_position().reshape(-1))
    return m


def mask_v2(env, rot, size):
    obs = env.cur_observation.astype(np.float32)      # what VecPyTorch hands to main.py:163-168
    if rot:
        import torch
        return np.asarray(REF.utils.get_rotation_mask(torch.from_numpy(obs), size), dtype=np.int32)
    return np.asarray(REF.utils.get_possible_position(obs, size), dtype=np.int32)


def rollout(tag, size, rot, n_bins, n_steps, traj_source, seed):
    L, W, H = size
    A = L * W
    R = 2 if rot else 1
    rng = np.random.default_rng(seed)
    # enough trajectories per bin for n_steps steps even if every step terminated an episode
    n_eps = n_steps + 1
    trajs = [[traj_source(rng) for _ in range(n_eps)] for _ in range(n_bins)]
    with rl.quiet():
        envs = [REF.PackingGame(box_creator=ScriptedBoxCreator(trajs[b]), container_size=size,
                                enable_rotation=rot) for b in range(n_bins)]
    box = np.zeros((n_steps + 1, n_bins, 3), np.uint8)
    hm = np.zeros((n_steps + 1, n_bins, A), np.uint8)
    m1 = np.zeros((n_steps + 1, n_bins, A * R), np.uint8)
    m2 = np.zeros((n_steps + 1, n_bins, A * R), np.uint8)
    act = np.zeros((n_steps, n_bins), np.int64)
    rew = np.zeros((n_steps, n_bins), np.float64)
    done = np.zeros((n_steps, n_bins), np.uint8)
    cnt = np.zeros((n_steps, n_bins), np.int32)
    rat = np.zeros((n_steps, n_bins), np.float64)
    hm_term = np.zeros((n_steps, n_bins, A), np.uint8)   # heightmap right after the step, before auto-reset

    def record(t, b, env):
        box[t, b] = env.next_box
        hm[t, b] = env.space.plain.reshape(-1)
        m1[t, b] = mask_v1(env, rot)
        m2[t, b] = mask_v2(env, rot, size)

    for b, env in enumerate(envs):
        env.reset()
        record(0, b, env)
    for t in range(n_steps):
        for b, env in enumerate(envs):
            u = rng.random()
            valid = np.flatnonzero(m2[t, b])
            if u < 0.88:
                a = int(rng.choice(valid))
            elif u < 0.96:
                a = int(rng.integers(0, A * R))
            else:   # edge actions; without rotation anything > A trips the reference's assert (bin3D.py:103)
                a = int(rng.choice([A, -1, 0, A - 1] + ([A * R - 1, A * R, A * R + 7, A + 1] if rot else [])))
            act[t, b] = a
            _, r, d, info = env.step([a])
            rew[t, b], done[t, b], cnt[t, b], rat[t, b] = r, d, info["counter"], info["ratio"]
            hm_term[t, b] = env.space.plain.reshape(-1)
            if d:
                env.reset()
            record(t + 1, b, env)
    # pack the injected trajectories the way the batched replay generator wants them:
    # row (b + e * n_bins) is episode e of bin b  (first_traj = 0)
    K = max(len(tr) for per_bin in trajs for tr in per_bin)
    T = n_bins * n_eps
    traj = np.zeros((T, K, 3), np.uint8)
    traj_len = np.zeros(T, np.int32)
    for b in range(n_bins):
        for e in range(n_eps):
            tr = trajs[b][e]
            traj[b + e * n_bins, :len(tr)] = np.asarray(tr, np.uint8).reshape(-1, 3)
            traj_len[b + e * n_bins] = len(tr)
    out = os.path.join(HERE, "ref_rollouts_%s.npz" % tag)
    np.savez_compressed(out, size=np.array(size, np.int32), rot=np.int32(rot), box=box, hm=hm, mask_v1=m1,
                        mask_v2=m2, action=act, reward=rew, done=done, counter=cnt, ratio=rat,
                        hm_term=hm_term, traj=traj, traj_len=traj_len)
    print("%s: %d bins x %d steps, %d episodes ended, v1!=v2 on %d masks, %d bytes" % (
        tag, n_bins, n_steps, int(done.sum()), int((m1 != m2).any(axis=2).sum()), os.path.getsize(out)))


def load_pt(name):
    import torch
    return torch.load(os.path.join(REF.path, "dataset", name))


def main():
    np.random.seed(0)
    random.seed(0)
    cut2 = load_pt("cut_2.pt")
    cut1 = load_pt("cut_1.pt")
    rs = load_pt("rs.pt")

    def from_cut2(rng):
        return [list(b) for b in cut2[int(rng.integers(0, len(cut2)))]]

    def from_rs_short(rng):   # short rs-style sequences so that the sentinel path is exercised too
        n = int(rng.integers(1, 12))
        return [[int(v) for v in rng.integers(2, 6, 3)] for _ in range(n)]

    def mixed10(rng):
        return from_cut2(rng) if rng.random() < 0.6 else from_rs_short(rng)

    def big20(rng):           # items up to 10x10 footprints reach the v1 != v2 region (area > 40)
        n = int(rng.integers(5, 60))
        out = []
        for _ in range(n):
            if rng.random() < 0.35:
                out.append([int(rng.integers(2, 11)), int(rng.integers(2, 11)), int(rng.integers(1, 5))])
            else:
                out.append([int(v) for v in rng.integers(2, 6, 3)])
        return out

    def odd(rng):
        n = int(rng.integers(3, 30))
        return [[int(rng.integers(1, 6)), int(rng.integers(1, 7)), int(rng.integers(1, 5))] for _ in range(n)]

    rollout("10_norot", (10, 10, 10), False, 48, 48, mixed10, 11)
    rollout("10_rot", (10, 10, 10), True, 48, 48, mixed10, 12)
    rollout("20_norot", (20, 20, 20), False, 12, 96, big20, 13)
    rollout("8x12x9_rot", (8, 12, 9), True, 24, 48, odd, 14)

    # --- known-answer data of the reference repo ---------------------------------------------------
    rows = [ln.strip() for ln in open(os.path.join(REF.path, "user_study", "trajs", "traj0.txt"))][2:]
    placed = np.array([[int(v) for v in r.split(",")] for r in rows[:-1]], np.int32)   # x,y,z,lx,ly,lz
    last = np.array([int(v) for v in rows[-1].split(",")], np.int32)
    np.savez_compressed(os.path.join(HERE, "traj0.npz"), placed=placed, last=last)

    trajs, cur = [], None
    with open(os.path.join(REF.path, "user_study", "cut_2.txt")) as f:
        for ln in f:
            if ln.startswith("The"):
                if len(trajs) == 300:
                    break
                cur = []
                trajs.append(cur)
            elif ln.strip():
                cur.append([int(v) for v in ln.split()])
    K = max(len(t) for t in trajs)
    arr = np.zeros((len(trajs), K, 6), np.uint8)
    ln_ = np.zeros(len(trajs), np.int32)
    for i, t in enumerate(trajs):
        arr[i, :len(t)] = np.asarray(t, np.uint8)
        ln_[i] = len(t)
    np.savez_compressed(os.path.join(HERE, "cut2_txt_sample.npz"), pieces=arr, n=ln_)

    def pad(ds):
        K = max(len(t) for t in ds)
        a = np.zeros((len(ds), K, 3), np.uint8)
        n = np.zeros(len(ds), np.int32)
        for i, t in enumerate(ds):
            a[i, :len(t)] = np.asarray(t, np.uint8).reshape(-1, 3)
            n[i] = len(t)
        return a, n
    c1, n1 = pad(cut1)
    c2, n2 = pad(cut2)
    r_, nr = pad(rs)
    np.savez_compressed(os.path.join(HERE, "datasets.npz"), cut_1=c1, cut_1_len=n1, cut_2=c2, cut_2_len=n2,
                        rs=r_, rs_len=nr)

    # --- generator statistics (reference generators, MT19937) --------------------------------------
    stats = {}
    for size, n_ep in (((10, 10, 10), 4000), ((20, 20, 20), 200)):
        tag = "%d" % size[0]
        for kind in ("cut1", "cut1rot", "cut2"):
            with rl.quiet():
                if kind == "cut2":
                    gen = REF.MDlayerBoxCreator(size, [2, 5])
                else:
                    gen = REF.CuttingBoxCreator(size, [2, 2, 2, 5, 5, 5], kind == "cut1rot")
            lens = np.zeros(n_ep, np.int32)
            dims = np.zeros((3, 16), np.int64)          # histogram of item dims per axis
            first = np.zeros((3, 16), np.int64)         # same for the first item of each episode
            zsum_prefix = np.zeros(64, np.float64)      # mean z of the k-th item (order statistics)
            zcnt_prefix = np.zeros(64, np.int64)
            for e in range(n_ep):
                with rl.quiet():
                    gen.reset()
                seq = []
                if kind == "cut2":
                    seq = [list(b) for b in gen.box_set[:-1]]
                else:
                    while len(gen.candidates) > 0:     # empty candidates => sentinel (cutCreator.py:113-115)
                        gen.box_list.clear()
                        gen.generate_box_size()
                        seq.append(list(gen.box_list[0]))
                assert sum(x * y * z for x, y, z in seq) == size[0] * size[1] * size[2]
                lens[e] = len(seq)
                for k, (x, y, z) in enumerate(seq):
                    dims[0, x] += 1; dims[1, y] += 1; dims[2, z] += 1
                    if k < 64:
                        zsum_prefix[k] += z; zcnt_prefix[k] += 1
                x, y, z = seq[0]
                first[0, x] += 1; first[1, y] += 1; first[2, z] += 1
            stats["%s_%s_len" % (kind, tag)] = lens
            stats["%s_%s_dims" % (kind, tag)] = dims
            stats["%s_%s_first" % (kind, tag)] = first
            stats["%s_%s_zk" % (kind, tag)] = zsum_prefix / np.maximum(zcnt_prefix, 1)
            print(kind, tag, "len mean %.2f min %d max %d" % (lens.mean(), lens.min(), lens.max()))
    np.savez_compressed(os.path.join(HERE, "gen_stats.npz"), **stats)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
