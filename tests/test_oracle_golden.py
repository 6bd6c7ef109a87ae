"""Pins the CPU oracle (oracle/bpp_oracle.c) against fixtures recorded from the Python reference
(tests/golden/make_golden.py) and against the reference repo's own known-answer data."""
import os

import numpy as np
import pytest

import oracle as orc

ROLLOUTS = ["10_norot", "10_rot", "20_norot", "8x12x9_rot"]


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_threshold_integer_equivalence():
    # max_area/area > 0.95|0.85|0.50 in float64 (space.py:137-141) == the integer forms of the oracle
    for a in range(1, 4097):
        ma = np.arange(0, a + 1)
        q = ma / a
        assert np.array_equal(q > 0.95, 20 * ma > 19 * a)
        assert np.array_equal(q > 0.85, 20 * ma > 17 * a)
        assert np.array_equal(q > 0.50, 2 * ma > a)


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32_10
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        assert tuple(int(v) for v in orc.philox(ctr, key)) == want
    # stream layout: n-th draw = word n&3 of block n>>2 on the (gid) subsequence
    for n in (0, 1, 5, 1023):
        blk = orc.philox((n >> 2, 0, 77, 0), (1234, 0))
        assert orc.draw(1234, 0, 0, 77, n) == int(blk[n & 3])


@pytest.mark.parametrize("tag", ROLLOUTS)
def test_masks_match_reference(golden_dir, tag):
    g = load(golden_dir, "ref_rollouts_%s.npz" % tag)
    size, rot = tuple(int(v) for v in g["size"]), bool(g["rot"])
    T1, N = g["hm"].shape[:2]
    for t in range(T1):
        for b in range(N):
            plain, box = g["hm"][t, b].astype(np.int32), g["box"][t, b]
            assert np.array_equal(orc.mask(plain, box, size, rot, orc.MASK_V2), g["mask_v2"][t, b]), (t, b)
            assert np.array_equal(orc.mask(plain, box, size, rot, orc.MASK_V1), g["mask_v1"][t, b]), (t, b)


@pytest.mark.parametrize("tag", ROLLOUTS)
@pytest.mark.parametrize("variant", [orc.MASK_V1, orc.MASK_V2])
def test_env_rollout_matches_reference(golden_dir, tag, variant):
    g = load(golden_dir, "ref_rollouts_%s.npz" % tag)
    size, rot = tuple(int(v) for v in g["size"]), bool(g["rot"])
    T, N = g["action"].shape
    env = orc.OracleEnv(N, size, rot=rot, mask_variant=variant, gen="replay", traj=g["traj"],
                        traj_len=g["traj_len"], first_traj=0, n_threads=2)
    A = size[0] * size[1]
    want_mask = g["mask_v1"] if variant == orc.MASK_V1 else g["mask_v2"]

    def check_obs(t):
        assert np.array_equal(env.obs[:, :A], g["hm"][t].astype(np.float32))
        for c in range(3):
            assert np.array_equal(env.obs[:, (c + 1) * A:(c + 2) * A],
                                  np.repeat(g["box"][t][:, c:c + 1], A, axis=1).astype(np.float32))
        assert np.array_equal(env.mask, want_mask[t].astype(np.float32))
        plain, box, _, _ = env.get_state()
        assert np.array_equal(plain, g["hm"][t].astype(np.int32))
        assert np.array_equal(box, g["box"][t].astype(np.int32))

    env.reset()
    check_obs(0)
    for t in range(T):
        env.step(g["action"][t])
        assert np.array_equal(env.done, g["done"][t]), t
        assert np.array_equal(env.counter, g["counter"][t]), t
        assert np.array_equal(env.reward, g["reward"][t].astype(np.float32)), t   # f64 -> f32 as acktr/envs.py:192
        assert np.max(np.abs(env.ratio.astype(np.float64) - g["ratio"][t])) <= 1e-6
        check_obs(t + 1)
    m = env.metrics()
    assert m["episodes"] == int(g["done"].sum())
    assert m["sum_counter"] == int(g["counter"][g["done"] == 1].sum())


def test_traj0_known_answer(golden_dir):
    # user_study/trajs/traj0.txt: 6 placements reproduce their lz, the 7th box has no feasible position
    g = load(golden_dir, "traj0.npz")
    plain = np.zeros((10, 10), np.int32)
    for x, y, z, lx, ly, lz in g["placed"]:
        for variant in (orc.MASK_V1, orc.MASK_V2):
            assert orc.check_box(plain, x, y, lx, ly, z, (10, 10, 10), variant) == lz
        plain[lx:lx + x, ly:ly + y] = lz + z
    vol = int(sum(x * y * z for x, y, z, *_ in g["placed"]))
    assert vol / 1000 == pytest.approx(0.3)
    x, y, z = (int(v) for v in g["last"])
    for variant in (orc.MASK_V1, orc.MASK_V2):
        for i in range(10):
            for j in range(10):
                assert orc.check_box(plain, x, y, i, j, z, (10, 10, 10), variant) == -1
        assert orc.mask(plain, (x, y, z), (10, 10, 10), False, variant).all()       # all-ones fallback


def test_cut2_txt_known_answers(golden_dir):
    # user_study/cut_2.txt: pieces with positions; replay in support order => check_box == lz everywhere
    g = load(golden_dir, "cut2_txt_sample.npz")
    for tr, n in zip(g["pieces"], g["n"]):
        rows = sorted((tuple(int(v) for v in r) for r in tr[:n]), key=lambda r: r[5])
        plain = np.zeros((10, 10), np.int32)
        for x, y, z, lx, ly, lz in rows:
            for variant in (orc.MASK_V1, orc.MASK_V2):
                assert orc.check_box(plain, x, y, lx, ly, z, (10, 10, 10), variant) == lz
            plain[lx:lx + x, ly:ly + y] = lz + z
        assert (plain == 10).all()


def test_v1_v2_divergence_examples():
    # 20x20 bin, 7x7 item on a plateau of 3 with two opposite corners at 0, z = 2
    plain = np.zeros((20, 20), np.int32)
    plain[:7, :7] = 3
    plain[0, 0] = 0
    plain[6, 6] = 0
    assert orc.check_box(plain, 7, 7, 0, 0, 2, (20, 20, 20), orc.MASK_V1) == -1
    assert orc.check_box(plain, 7, 7, 0, 0, 2, (20, 20, 20), orc.MASK_V2) == 3
    # 10x10 window at height 3 with corners (0,1,1,0)
    plain = np.full((10, 10), 3, np.int32)
    plain[0, 0], plain[9, 0], plain[0, 9], plain[9, 9] = 0, 1, 1, 0
    assert orc.check_box(plain, 10, 10, 0, 0, 1, (10, 10, 10), orc.MASK_V1) == -1
    assert orc.check_box(plain, 10, 10, 0, 0, 1, (10, 10, 10), orc.MASK_V2) == 3


def test_replay_dataset_semantics(golden_dir):
    # binCreator.py:51-72: first episode plays trajectory 1; past the stored row -> (10,10,10) forever
    d = load(golden_dir, "datasets.npz")
    env = orc.OracleEnv(3, (10, 10, 10), gen="replay", traj=d["cut_2"], traj_len=d["cut_2_len"], first_traj=1)
    env.reset()
    _, box, _, _ = env.get_state()
    for b in range(3):
        assert tuple(box[b]) == tuple(int(v) for v in d["cut_2"][1 + b, 0])
    # every cut_2 row already ends with the stored [10,10,10]
    last = d["cut_2"][np.arange(len(d["cut_2"])), d["cut_2_len"] - 1]
    assert (last == 10).all()


def test_edge_actions_and_sizes():
    # item larger than the bin -> no position -> all-ones mask; any action then fails -> done
    traj = np.zeros((4, 2, 3), np.uint8)
    traj[:, 0] = (12, 3, 2)
    traj[:, 1] = (3, 3, 3)
    env = orc.OracleEnv(1, (10, 10, 10), gen="replay", traj=traj, traj_len=np.full(4, 2, np.int32), first_traj=0)
    env.reset()
    assert env.mask.all()
    env.step([0])
    assert env.done[0] == 1 and env.reward[0] == 0 and env.counter[0] == 0
    # action == A (un-rotated, out of bounds: bin3D.py:102 uses '>') and negative actions terminate too
    env2 = orc.OracleEnv(2, (10, 10, 10), rot=True, gen="rs", seed=3)
    env2.reset()
    env2.step([100, -5])
    assert env2.done.tolist() == [1, 1]


def test_f32_division_equals_reference_route():
    """The CUDA kernel computes reward = fdiv_rn(float(10*v), float(LWH)) and ratio = fdiv_rn(float(vol), float(LWH));
    the reference computes float32(float64(v)/float64(LWH)*10.0) and float32(float64(vol)/float64(LWH))
    (bin3D.py:44-46,121, space.py:146-151, acktr/envs.py:192).  Both routes give the same float32."""
    rng = np.random.default_rng(0)
    sizes = [1000, 8000, 864, 378, 125, 27000, 32 * 32 * 255, 7 * 11 * 13, 1]
    sizes += [int(a) * int(b) * int(c) for a, b, c in rng.integers(1, 33, (40, 3))]
    for lwh in sizes:
        v = np.arange(0, lwh + 1, dtype=np.int64)
        if len(v) > 300000:
            v = np.unique(np.concatenate([v[:50000], rng.integers(0, lwh + 1, 200000)]))
        ref_reward = ((v.astype(np.float64) / np.float64(lwh)) * 10.0).astype(np.float32)
        ref_ratio = (v.astype(np.float64) / np.float64(lwh)).astype(np.float32)
        ok = 10 * v < 2 ** 24
        got_reward = (10 * v[ok]).astype(np.float32) / np.float32(lwh)
        got_ratio = v.astype(np.float32) / np.float32(lwh)
        assert got_reward.dtype == np.float32 and np.array_equal(got_reward, ref_reward[ok]), lwh
        assert np.array_equal(got_ratio, ref_ratio), lwh


def test_rules_v1_and_v2_agree_on_small_footprints():
    """For footprints of at most 40 cells rule v1 (space.py:111-144) and rule v2 (acktr/utils.py:8-35) give the same verdict --
    which is why the kernel's packed sweep of the 16 unrolled footprints (sides 2..5) serves both.  Reachable and noisy maps."""
    import oracle as orc
    rng = np.random.default_rng(0)
    for t in range(20000):
        L = W = 10
        if t & 1:
            plain = rng.integers(0, 4, (L, W)).astype(np.int32)
        else:
            plain = np.zeros((L, W), np.int32)
            for _ in range(rng.integers(0, 10)):
                x, y = rng.integers(1, 7, 2)
                i, j = rng.integers(0, L - x + 1), rng.integers(0, W - y + 1)
                plain[i:i + x, j:j + y] = min(10, plain[i:i + x, j:j + y].max() + rng.integers(0, 4))
        x, y = (rng.integers(2, 6, 2) if t % 3 else rng.integers(1, 9, 2))
        if x * y > 40:
            continue
        z = int(rng.integers(1, 6))
        lx, ly = rng.integers(0, L - x + 1), rng.integers(0, W - y + 1)
        assert orc.check_box(plain, x, y, lx, ly, z, (L, W, 10), 1) == orc.check_box(plain, x, y, lx, ly, z, (L, W, 10), 2)
