"""Item-sequence generators of the oracle (Philox-driven restatements of cutCreator.py / mdCreator.py /
RandomBoxCreator) against invariants and against statistics recorded from the reference's own
(MT19937-driven) generators and datasets (tests/golden/gen_stats.npz, datasets.npz).

The reference's own MT19937 streams cannot be reproduced, so these tests pin the *distribution*; the algorithm itself
is pinned episode for episode in tests/test_generator_identity.py (reference generators fed with the Philox stream)."""
import os

import numpy as np
import pytest

import oracle as orc

LO, HI = (2, 2, 2), (5, 5, 5)


def episodes(kind, size, n, rot=False, seed=7):
    return [orc.gen_episode(kind, size, LO, HI, seed, gid, ep, rot=rot) for gid in range(n // 4) for ep in range(4)]


def check_partition(items, pos, size):
    L, W, H = size
    occ = np.zeros(size, np.int32)
    for (x, y, z), (lx, ly, lz) in zip(items.astype(int), pos.astype(int)):
        assert lx + x <= L and ly + y <= W and lz + z <= H
        assert (occ[lx:lx + x, ly:ly + y, lz:lz + z] == 0).all()
        occ[lx:lx + x, ly:ly + y, lz:lz + z] = 1
    assert occ.all()           # exact partition of the bin (dataset/cut_1.pt has the same property)


@pytest.mark.parametrize("kind", ["cut1", "cut2"])
@pytest.mark.parametrize("size", [(10, 10, 10), (20, 20, 20)])
def test_cut_invariants(kind, size):
    for items, pos in episodes(kind, size, 40 if size[0] == 10 else 8):
        assert ((items >= 2) & (items <= 5)).all()
        assert int(items.astype(np.int64).prod(1).sum()) == size[0] * size[1] * size[2]
        check_partition(items, pos, size)
        if kind == "cut2":     # stable sort by low_bound, mdCreator.py:137-138
            assert (np.diff(pos[:, 2].astype(int)) >= 0).all()
        else:                  # support order: every piece rests exactly on what was emitted before (cutCreator.py:98-107)
            plain = np.zeros(size[:2], np.int32)
            for (x, y, z), (lx, ly, lz) in zip(items.astype(int), pos.astype(int)):
                assert (plain[lx:lx + x, ly:ly + y] == lz).all()
                plain[lx:lx + x, ly:ly + y] += z
            assert (plain == size[2]).all()


def test_cut1_rotation_keeps_volume_and_lengths(golden_dir):
    # with rotation each emitted item is (x, y, z) or (y, x, z) with probability 1/2 (cutCreator.py:120-126)
    g = np.load(os.path.join(golden_dir, "gen_stats.npz"))
    lens = []
    for gid in range(400):
        items, _ = orc.gen_episode("cut1", (10, 10, 10), LO, HI, 3, gid, 0, rot=True)
        assert ((items >= 2) & (items <= 5)).all()
        assert int(items.astype(np.int64).prod(1).sum()) == 1000
        lens.append(len(items))
    ref = g["cut1rot_10_len"]
    assert abs(np.mean(lens) - ref.mean()) < 1.0
    # x and y marginals become exchangeable
    dims = np.zeros((2, 16))
    for gid in range(400):
        items, _ = orc.gen_episode("cut1", (10, 10, 10), LO, HI, 3, gid, 0, rot=True)
        for a in range(2):
            dims[a] += np.bincount(items[:, a], minlength=16)
    p = dims / dims.sum(1, keepdims=True)
    assert np.abs(p[0] - p[1]).max() < 0.03


@pytest.mark.parametrize("kind,tag", [("cut1", "cut1_10"), ("cut2", "cut2_10"), ("cut1", "cut1_20"), ("cut2", "cut2_20")])
def test_cut_distribution_matches_reference(golden_dir, kind, tag):
    g = np.load(os.path.join(golden_dir, "gen_stats.npz"))
    size = (10, 10, 10) if tag.endswith("10") else (20, 20, 20)
    n = 2000 if size[0] == 10 else 120
    eps = [orc.gen_episode(kind, size, LO, HI, 11, gid, 0)[0] for gid in range(n)]
    lens = np.array([len(e) for e in eps])
    ref = g[tag + "_len"]
    se = np.sqrt(lens.var() / len(lens) + ref.var() / len(ref))
    assert abs(lens.mean() - ref.mean()) < 5 * se + 0.02 * ref.mean(), (lens.mean(), ref.mean())
    assert abs(lens.std() - ref.std()) < 0.25 * ref.std()
    # per-axis marginals of item dimensions
    dims = np.zeros((3, 16))
    for e in eps:
        for a in range(3):
            dims[a] += np.bincount(e[:, a], minlength=16)
    p, q = dims / dims.sum(1, keepdims=True), g[tag + "_dims"] / g[tag + "_dims"].sum(1, keepdims=True)
    assert np.abs(p - q).max() < 0.03, np.abs(p - q).max()
    # layer structure: mean z of the k-th emitted item
    k = 10
    zk = np.array([np.mean([e[i, 2] for e in eps if len(e) > i]) for i in range(k)])
    assert np.abs(zk - g[tag + "_zk"][:k]).max() < 0.35


def test_cut_matches_datasets(golden_dir):
    # dataset/cut_1.pt and cut_2.pt were produced by the reference generators: same length / dim statistics
    d = np.load(os.path.join(golden_dir, "datasets.npz"))
    for kind, key, extra in (("cut1", "cut_1", 0), ("cut2", "cut_2", 1)):
        ref_len = d[key + "_len"] - extra           # cut_2 rows carry the stored [10,10,10] sentinel
        lens = np.array([len(orc.gen_episode(kind, (10, 10, 10), LO, HI, 5, gid, 1)[0]) for gid in range(1500)])
        assert abs(lens.mean() - ref_len.mean()) < 1.0, (kind, lens.mean(), ref_len.mean())
        assert lens.min() >= ref_len.min() - 4 and lens.max() <= ref_len.max() + 8


def test_rs_is_uniform_over_item_set():
    env = orc.OracleEnv(4096, (10, 10, 10), gen="rs", seed=99)
    env.reset()
    counts = np.zeros((3, 6))
    for _ in range(8):
        _, box, _, _ = env.get_state()
        for a in range(3):
            counts[a] += np.bincount(box[:, a], minlength=6)
        env.step(env.sample.copy())
    p = counts[:, 2:] / counts[:, 2:].sum(1, keepdims=True)
    assert np.abs(p - 0.25).max() < 0.01            # each of 2..5 with probability 1/4 (np.random.randint over 64 items)


def test_env_with_cut_generators_runs_and_partitions():
    for kind in ("cut1", "cut2"):
        env = orc.OracleEnv(64, (10, 10, 10), gen=kind, seed=5)
        env.reset()
        for _ in range(40):
            env.step(env.sample.copy())
        m = env.metrics()
        assert m["episodes"] > 0 and m["sum_vol"] <= m["episodes"] * 1000
