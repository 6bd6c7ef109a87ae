"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures recorded
from the Python reference.  Integer outputs bit-exact; reward bit-exact in float32; ratio within 1e-6."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle as orc

pytestmark = pytest.mark.gpu

ROLLOUTS = ["10_norot", "10_rot", "20_norot", "8x12x9_rot"]


@pytest.fixture(scope="module")
def env_mod():
    import torch
    import b200bpp
    assert torch.cuda.is_available()
    return b200bpp


def t2n(t):
    return t.detach().cpu().numpy()


def test_philox_matches_oracle_and_curand(env_mod):
    import torch
    from b200bpp import _cabi
    l = _cabi.lib()
    n = 64
    out = torch.zeros(n, dtype=torch.int32, device="cuda")
    for seed, dom, aux, gid in [(1234, 0, 0, 0), (1234, 0, 0, 77), (0xDEADBEEFCAFE, 1, 0, 2 ** 33 + 5), (7, 2, 12345, 9)]:
        _cabi.check(l.bpp_debug_draws(seed, dom, aux, gid, n, C.c_void_p(out.data_ptr()), None))
        got = t2n(out).view(np.uint32)
        want = np.array([orc.draw(seed, dom, aux, gid, i) for i in range(n)], dtype=np.uint32)
        assert np.array_equal(got, want)
        if dom == 0 and aux == 0:   # the item stream is cuRAND's Philox4_32_10 stream (curand_init(seed, gid, 0))
            ref = torch.zeros(n, dtype=torch.int32, device="cuda")
            _cabi.check(l.bpp_debug_curand_draws(seed, gid, n, C.c_void_p(ref.data_ptr()), None))
            assert np.array_equal(t2n(ref).view(np.uint32), want)


@pytest.mark.parametrize("tag", ROLLOUTS)
@pytest.mark.parametrize("variant", [1, 2])
def test_golden_rollout(env_mod, golden_dir, tag, variant):
    """Replays the traces recorded from the Python reference through the CUDA env."""
    g = np.load(os.path.join(golden_dir, "ref_rollouts_%s.npz" % tag))
    size, rot = tuple(int(v) for v in g["size"]), bool(g["rot"])
    T, N = g["action"].shape
    A = size[0] * size[1]
    env = env_mod.BinPackingVecEnv(N, size, enable_rotation=rot, trajectories=(g["traj"], g["traj_len"]),
                                   first_traj=0, mask_variant=variant)
    want_mask = g["mask_v1"] if variant == 1 else g["mask_v2"]

    def check(t):
        obs = t2n(env.obs)
        assert np.array_equal(obs[:, :A], g["hm"][t].astype(np.float32)), t
        for c in range(3):
            assert np.array_equal(obs[:, (c + 1) * A:(c + 2) * A], np.repeat(g["box"][t][:, c:c + 1], A, 1).astype(np.float32))
        assert np.array_equal(t2n(env.location_masks), want_mask[t].astype(np.float32)), t
        plain, box, _, _ = env.export_state()
        assert np.array_equal(t2n(plain), g["hm"][t].astype(np.int32))
        assert np.array_equal(t2n(box), g["box"][t].astype(np.int32))

    env.reset()
    check(0)
    for t in range(T):
        if t % 2 == 0:     # alternate the device-resident and the host-buffer entry points
            import torch
            env.step_device(torch.as_tensor(g["action"][t], device="cuda"))
            reward, done, counter, ratio = t2n(env.reward_t), t2n(env.done_t), t2n(env.counter_t), t2n(env.ratio_t)
        else:
            _, r, d, infos = env.step(g["action"][t].reshape(-1, 1))
            reward, done = r.numpy().reshape(-1), d.astype(np.uint8)
            counter = np.array([infos[i]["counter"] for i in range(N)])
            ratio = np.array([infos[i]["ratio"] for i in range(N)], dtype=np.float32)
        assert np.array_equal(done, g["done"][t]), t
        assert np.array_equal(counter, g["counter"][t]), t
        assert np.array_equal(reward, g["reward"][t].astype(np.float32)), t
        assert np.max(np.abs(ratio.astype(np.float64) - g["ratio"][t])) <= 1e-6
        check(t + 1)
    m = env.metrics()
    assert m["episodes"] == int(g["done"].sum())
    assert m["sum_counter"] == int(g["counter"][g["done"] == 1].sum())
    assert m["steps"] == T * N
    env.close()


def run_against_oracle(env_mod, N, size, rot, variant, gen_kw, T, seed, offset=0, global_n=None, edge_every=0):
    """Steps the CUDA env and the oracle with the device-sampled actions; compares everything."""
    import torch
    okw = dict(gen_kw)
    env = env_mod.BinPackingVecEnv(N, size, enable_rotation=rot, mask_variant=variant, seed=seed,
                                   global_offset=offset, global_n=global_n, want_sample=True, **gen_kw)
    if "trajectories" in okw:
        tr = okw.pop("trajectories")
        okw.pop("first_traj", None)
        okw.update(gen="replay", traj=tr[0], traj_len=tr[1])
    else:
        okw = dict(gen=okw.get("data_type", "rs"))
        if "box_set" in gen_kw:
            okw["item_set"] = np.asarray(gen_kw["box_set"], np.uint8)
    o = orc.OracleEnv(N, size, rot=rot, mask_variant=variant, seed=seed, global_offset=offset, global_n=global_n,
                      first_traj=gen_kw.get("first_traj", 1), n_threads=0, **okw)
    env.reset()
    o.reset()
    A = size[0] * size[1]
    rng = np.random.default_rng(seed)
    for t in range(T + 1):
        assert np.array_equal(t2n(env.obs), o.obs), t
        assert np.array_equal(t2n(env.location_masks), o.mask), t
        assert np.array_equal(t2n(env.sample_t), o.sample), t
        if t == T:
            break
        actions = o.sample.copy()
        if edge_every and t % edge_every == 0:      # sprinkle out-of-range / quirk actions
            k = rng.integers(0, N, max(1, N // 16))
            actions[k] = rng.choice([A, -1, A * (2 if rot else 1) + 3, 0], size=k.shape[0]) if rot else rng.choice([A, -1, 0], size=k.shape[0])
        env.step_device(torch.as_tensor(actions, device="cuda"))
        o.step(actions)
        assert np.array_equal(t2n(env.done_t), o.done), t
        assert np.array_equal(t2n(env.counter_t), o.counter), t
        assert np.array_equal(t2n(env.reward_t), o.reward), t
        assert np.max(np.abs(t2n(env.ratio_t).astype(np.float64) - o.ratio)) <= 1e-6
    plain, box, counter, vol = env.export_state()
    oplain, obox, ocounter, ovol = o.get_state()
    assert np.array_equal(t2n(plain), oplain) and np.array_equal(t2n(box), obox)
    assert np.array_equal(t2n(counter), ocounter) and np.array_equal(t2n(vol), ovol)
    m, om = env.metrics(), o.metrics()
    assert m == om
    env.close()
    return m


@pytest.mark.parametrize("rot", [False, True])
@pytest.mark.parametrize("variant", [1, 2])
def test_config2_rs_4096_bins(env_mod, rot, variant):
    # BASELINE.json configs[1]: 4096 parallel bins, 10x10x10, rs sequence, mask bit-exact
    m = run_against_oracle(env_mod, 4096, (10, 10, 10), rot, variant, dict(data_type="rs"), T=64, seed=1234, edge_every=7)
    assert m["episodes"] > 4096


def test_config3_cut2_replay_subsample(env_mod, golden_dir):
    # configs[2]: cut_2.pt replayed; bin g, episode e plays trajectory (1 + g + e*N) mod 2100
    d = np.load(os.path.join(golden_dir, "datasets.npz"))
    run_against_oracle(env_mod, 1024, (10, 10, 10), False, 2, dict(trajectories=(d["cut_2"], d["cut_2_len"]), first_traj=1),
                       T=128, seed=5, offset=3000, global_n=65536)


def test_config4_20cube_subsample(env_mod):
    # configs[3]: 20x20x20 bin, rs; parity subsample 256 bins x 128 steps
    run_against_oracle(env_mod, 256, (20, 20, 20), False, 2, dict(data_type="rs"), T=128, seed=99, offset=1000, global_n=262144)


def test_config5_rotation_cut2_shard(env_mod, golden_dir):
    # configs[4]: rotation enabled, cut_2 replay, one shard of a 524288-bin job (rank 3 of 8)
    d = np.load(os.path.join(golden_dir, "datasets.npz"))
    run_against_oracle(env_mod, 512, (10, 10, 10), True, 2, dict(trajectories=(d["cut_2"], d["cut_2_len"]), first_traj=1),
                       T=96, seed=8, offset=3 * 65536, global_n=524288, edge_every=5)


@pytest.mark.parametrize("kind", ["cut1", "cut2"])
@pytest.mark.parametrize("rot", [False, True])
def test_cut_generators_on_device(env_mod, kind, rot):
    # episodes are cut on the device (bpp_gen.cu) with the oracle's Philox draw order
    run_against_oracle(env_mod, 777, (10, 10, 10), rot, 2, dict(data_type=kind), T=60, seed=31, offset=5000, global_n=100000, edge_every=9)


def test_cut_generators_20cube(env_mod):
    run_against_oracle(env_mod, 96, (20, 20, 20), False, 2, dict(data_type="cut2"), T=80, seed=32)
    run_against_oracle(env_mod, 96, (20, 20, 20), True, 1, dict(data_type="cut1"), T=80, seed=33)


def test_cut_explicit_second_reset(env_mod):
    # a second reset() moves every bin to its next episode (LoadBoxCreator-style index += 1), also for cut kinds
    env = env_mod.BinPackingVecEnv(64, data_type="cut2", seed=4, want_sample=True)
    o = orc.OracleEnv(64, gen="cut2", seed=4)
    for _ in range(3):
        env.reset(); o.reset()
        assert np.array_equal(t2n(env.obs), o.obs) and np.array_equal(t2n(env.location_masks), o.mask)
        for _ in range(5):
            a = o.sample.copy()
            import torch
            env.step_device(torch.as_tensor(a, device="cuda")); o.step(a)
            assert np.array_equal(t2n(env.obs), o.obs) and np.array_equal(t2n(env.done_t), o.done)


def test_odd_sizes_scalar_store_path(env_mod):
    # L*W not a multiple of 4 -> scalar stores; non-square; big item set incl. items larger than the bin
    items = [(x, y, z) for x in (1, 3, 8) for y in (2, 7) for z in (1, 4)]
    run_against_oracle(env_mod, 300, (7, 9, 6), True, 1, dict(data_type="rs", box_set=items), T=48, seed=3, edge_every=4)
    run_against_oracle(env_mod, 64, (5, 5, 5), False, 2, dict(data_type="rs", box_set=[(1, 1, 1), (2, 2, 2), (5, 5, 1)]), T=48, seed=4)


def test_sharding_is_transparent(env_mod):
    # two half-size envs with global offsets reproduce one full env (what multi-GPU sharding relies on)
    import torch
    N = 512
    full = env_mod.BinPackingVecEnv(N, seed=42, want_sample=True)
    halves = [env_mod.BinPackingVecEnv(N // 2, seed=42, want_sample=True, global_offset=k * N // 2, global_n=N) for k in range(2)]
    full.reset()
    [h.reset() for h in halves]
    for t in range(24):
        cat = lambda name: torch.cat([getattr(h, name) for h in halves])
        assert torch.equal(full.obs, cat("obs")) and torch.equal(full.location_masks, cat("location_masks"))
        assert torch.equal(full.sample_t, cat("sample_t"))
        full.step_device(full.sample_t.clone())
        [h.step_device(h.sample_t.clone()) for h in halves]
        assert torch.equal(full.reward_t, cat("reward_t")) and torch.equal(full.done_t, cat("done_t"))
    ms = [h.metrics() for h in halves]
    assert full.metrics()["episodes"] == sum(m["episodes"] for m in ms)
    assert full.metrics()["sum_vol"] == sum(m["sum_vol"] for m in ms)


def blocky_plains(rng, n, L, W, H):
    plains = np.zeros((n, L, W), np.int32)
    for p in plains:
        for _ in range(rng.integers(0, 14)):
            x, y = rng.integers(1, L + 1), rng.integers(1, W + 1)
            i, j = rng.integers(0, L - x + 1), rng.integers(0, W - y + 1)
            p[i:i + x, j:j + y] = min(H, p[i:i + x, j:j + y].max() + rng.integers(0, 4))
    return plains


@pytest.mark.parametrize("size", [(10, 10, 10), (20, 20, 20), (8, 12, 9)])
def test_stateless_masks_and_check_box_fuzz(env_mod, size):
    import torch
    from b200bpp import utils
    L, W, H = size
    rng = np.random.default_rng(L * 100 + W)
    n = 600
    plains = blocky_plains(rng, n, L, W, H)
    boxes = np.stack([rng.integers(1, min(L, 10) + 2, n), rng.integers(1, min(W, 10) + 2, n), rng.integers(1, 6, n)], 1).astype(np.int32)
    obs = np.concatenate([plains.reshape(n, -1)] + [np.repeat(boxes[:, c:c + 1], L * W, 1) for c in range(3)], 1).astype(np.float32)
    for rot in (False, True):
        for variant in (1, 2):
            want = np.stack([orc.mask(plains[i], boxes[i], size, rot, variant) for i in range(n)]).astype(np.float32)
            got = t2n(utils.mask_batch(torch.as_tensor(obs), size, rot, variant))
            assert np.array_equal(got, want)
            got2 = t2n(utils.mask_from_plain(plains.reshape(n, -1), boxes, size, rot, variant))
            assert np.array_equal(got2, want)
    q = np.stack([boxes[:, 0], boxes[:, 1], rng.integers(-1, L, n), rng.integers(-1, W, n), boxes[:, 2]], 1).astype(np.int32)
    for variant in (1, 2):
        want = np.array([orc.check_box(plains[i], *q[i][:5], size, variant) for i in range(n)], np.int32)
        assert np.array_equal(utils.check_box_batch(plains, q, size, variant), want)
    # reference-signature single-observation helpers (acktr/utils.py:37-94)
    assert utils.get_possible_position(torch.as_tensor(obs[0]).cuda(), size) == orc.mask(plains[0], boxes[0], size, False, 2).tolist()
    assert utils.get_possible_position(obs[1], size) == orc.mask(plains[1], boxes[1], size, False, 2).tolist()
    rm = utils.get_rotation_mask(torch.as_tensor(obs[2]), size)
    assert rm.dtype == np.int32 and np.array_equal(rm, orc.mask(plains[2], boxes[2], size, True, 2))


def test_v1_v2_divergence_on_device(env_mod):
    from b200bpp import utils
    plain = np.zeros((20, 20), np.int32)
    plain[:7, :7] = 3
    plain[0, 0] = plain[6, 6] = 0
    assert utils.check_box(plain, 7, 7, 0, 0, 2, (20, 20, 20), 1) == -1
    assert utils.check_box(plain, 7, 7, 0, 0, 2, (20, 20, 20), 2) == 3
    m1 = t2n(utils.mask_from_plain(plain.reshape(1, -1), [[7, 7, 2]], (20, 20, 20), False, 1))[0]
    m2 = t2n(utils.mask_from_plain(plain.reshape(1, -1), [[7, 7, 2]], (20, 20, 20), False, 2))[0]
    assert m1[0] == 0 and m2[0] == 1


def test_traj0_and_sentinel_on_device(env_mod, golden_dir):
    from b200bpp import utils
    g = np.load(os.path.join(golden_dir, "traj0.npz"))
    plain = np.zeros((10, 10), np.int32)
    for x, y, z, lx, ly, lz in g["placed"]:
        for variant in (1, 2):
            assert utils.check_box(plain, x, y, lx, ly, z, (10, 10, 10), variant) == lz
        plain[lx:lx + x, ly:ly + y] = lz + z
    m = t2n(utils.mask_from_plain(plain.reshape(1, -1), [g["last"]], (10, 10, 10), False, 2))
    assert m.all()      # no feasible position -> all-ones fallback


def test_vecenv_surface(env_mod):
    """What acktr/envs.py:77-118 hands to main.py: shapes, dtypes, devices, infos."""
    import torch
    env = env_mod.make_vec_envs("Bpp-v0", 1, 16, 0.99, None, "cuda:0", False,
                                args=type("A", (), dict(container_size=(10, 10, 10), enable_rotation=True,
                                                        box_size_set=env_mod.default_item_set(), data_type="rs")))
    assert env.num_envs == 16 and env.observation_space.shape == (400,) and env.action_space.n == 200
    assert env.action_space.__class__.__name__ == "Discrete"
    obs = env.reset()
    assert obs.shape == (16, 400) and obs.dtype == torch.float32 and obs.is_cuda
    assert env.location_masks.shape == (16, 200)
    a = torch.zeros((16, 1), dtype=torch.int64, device="cuda")
    a[0, 0] = 100      # action == A: un-rotated and out of bounds (bin3D.py:102) -> episode ends
    obs, reward, done, infos = env.step(a)
    assert reward.shape == (16, 1) and reward.dtype == torch.float32 and not reward.is_cuda
    assert done.dtype == np.bool_ and done.shape == (16,) and done[0]
    assert len(infos) == 16 and "episode" in infos[0] and infos[0]["episode"]["l"] == 1
    assert not done[1] and infos[1]["counter"] == 1 and abs(infos[1]["ratio"] - reward[1, 0].item() / 10) < 1e-6
    from b200bpp.utils import get_vec_normalize
    assert get_vec_normalize(env) is None
    env.close()


def test_errors_are_loud(env_mod):
    import torch
    from b200bpp import BppError
    env = env_mod.BinPackingVecEnv(8)
    with pytest.raises(BppError):
        env.step_device(torch.zeros(8, dtype=torch.int64, device="cuda"))     # step before reset
    env.reset()
    with pytest.raises(ValueError):
        env.step_device(torch.zeros(9, dtype=torch.int64, device="cuda"))
    with pytest.raises(BppError):
        env_mod.BinPackingVecEnv(8, container_size=(40, 10, 10))
    with pytest.raises(BppError):
        env_mod.BinPackingVecEnv(8, data_type="cut1", box_set=[(3, 3, 3), (4, 4, 4)])     # hi + 1 < 2 * lo: the reference asserts


def test_full_size_properties(env_mod):
    """BASELINE sizes through size-independent properties; a prefix of the bins is also checked against the oracle."""
    import torch
    for N, size, T in ((65536, (10, 10, 10), 48), (262144, (20, 20, 20), 24)):
        L, W, H = size
        A = L * W
        env = env_mod.BinPackingVecEnv(N, size, seed=77, want_sample=True)
        K = 128
        o = orc.OracleEnv(K, size, seed=77, global_n=N, n_threads=0)
        env.reset()
        o.reset()
        ep = 0
        vol_done = 0
        for t in range(T):
            assert torch.equal(env.obs[:K].cpu(), torch.from_numpy(o.obs))
            assert torch.equal(env.location_masks[:K].cpu(), torch.from_numpy(o.mask))
            mask = env.location_masks
            assert bool(((mask == 0) | (mask == 1)).all()) and bool((mask.sum(1) >= 1).all())
            assert bool((env.obs[:, :A] <= H).all()) and bool((env.obs[:, :A] >= 0).all())
            for c in range(1, 4):    # box planes are constant
                plane = env.obs[:, c * A:(c + 1) * A]
                assert bool((plane == plane[:, :1]).all())
            # the sampled action is valid under the mask
            assert bool((mask.gather(1, env.sample_t.view(-1, 1)) == 1).all())
            actions = env.sample_t.clone()
            env.step_device(actions)
            o.step(t2n(actions[:K]))
            assert np.array_equal(t2n(env.done_t[:K]), o.done) and np.array_equal(t2n(env.reward_t[:K]), o.reward)
            ep += int(env.done_t.sum())
            vol_done += int((env.ratio_t.double() * (L * W * H)).round().long()[env.done_t.bool()].sum())
            # reward is 10 * item volume / LWH exactly when not done, 0 when done
            assert bool((env.reward_t[env.done_t.bool()] == 0).all())
        plain, box, counter, vol = env.export_state()
        assert bool((plain.sum(1) >= vol).all())            # placed volume never exceeds the volume under the heightmap
        assert bool((plain.max(1).values <= H).all())
        m = env.metrics()
        assert m["episodes"] == ep and m["steps"] == N * T and m["sum_vol"] == vol_done
        env.close()


def test_single_env_view_matches_golden(env_mod, golden_dir):
    """PackingGame view (bin3D.py surface) replays a recorded reference trajectory: obs, reward, done, info, mask."""
    import copy
    from b200bpp.bpp0 import PackingGame, BoxCreator
    g = np.load(os.path.join(golden_dir, "ref_rollouts_10_rot.npz"))
    N = g["action"].shape[1]

    class Scripted(BoxCreator):
        def __init__(self, rows):
            super().__init__()
            self.rows, self.index, self.k = rows, -1, 0

        def reset(self):
            self.box_list.clear(); self.index += 1; self.k = 0

        def generate_box_size(self, **kw):
            row, n = self.rows[self.index]
            self.box_list.append(tuple(int(v) for v in row[self.k]) if self.k < n else (10, 10, 10))
            self.k += 1

    for b in (0, 7):
        rows = [(g["traj"][b + e * N], int(g["traj_len"][b + e * N])) for e in range(g["action"].shape[0] + 1)]
        env = PackingGame(box_creator=Scripted(rows), container_size=(10, 10, 10), enable_rotation=True)
        obs = env.reset()
        for t in range(20):
            assert np.array_equal(obs[:100], g["hm"][t, b]) and tuple(env.next_box) == tuple(g["box"][t, b])
            clone = copy.deepcopy(env)                      # reorder.py:69,226,251 deep-copies the env
            obs, r, d, info = env.step([int(g["action"][t, b])])
            assert d == bool(g["done"][t, b]) and abs(r - g["reward"][t, b]) < 1e-9
            assert info["counter"] == g["counter"][t, b] and abs(info["ratio"] - g["ratio"][t, b]) < 1e-9
            assert np.array_equal(clone.space.plain.reshape(-1), g["hm"][t, b])   # the copy kept the old state
            if d:
                obs = env.reset()
        env.close()
    # env-side mask (v1) of the view == golden mask_v1 (no rotation)
    g2 = np.load(os.path.join(golden_dir, "ref_rollouts_10_norot.npz"))
    rows = [(g2["traj"][e * g2["action"].shape[1]], int(g2["traj_len"][e * g2["action"].shape[1]])) for e in range(4)]
    env = PackingGame(box_creator=Scripted(rows), container_size=(10, 10, 10))
    env.reset()
    assert np.array_equal(env.get_possible_position().reshape(-1), g2["mask_v1"][0, 0])
    assert env.space.get_action_space() == 100 and env.space.position_to_index((3, 4)) == 34
    env.close()


def test_device_rollout_glue(env_mod):
    """main.py:148-174 on the device: storage shapes as acktr/storage.py, stats as Monitor, everything on the GPU."""
    import torch
    from b200bpp import rollout
    env = env_mod.BinPackingVecEnv(2048, (10, 10, 10), enable_rotation=True, seed=6)
    st = rollout.start(env, rollout.DeviceRolloutStorage(5, env))
    stats = rollout.EpisodeStats(env)
    torch.manual_seed(0)
    for _ in range(6):
        rollout.collect(env, st, stats=stats)
        # every stored action was valid under the stored mask, every stored mask row has a feasible entry
        assert bool((st.location_masks[:5].gather(2, st.actions) == 1).all())
        assert bool((st.location_masks.sum(2) >= 1).all())
        # masks[t+1] == 0 exactly where reward == 0 at t (a failed placement ends the episode)
        assert torch.equal(st.masks[1:] == 0, st.rewards == 0)
        st.after_update()
    s = stats.read()
    m = env.metrics()
    assert s["episodes"] == m["episodes"] and abs(s["mean_ratio"] - m["sum_vol"] / m["episodes"] / 1000) < 1e-6
    assert abs(s["mean_length"] - (m["sum_counter"] / m["episodes"] + 1)) < 1e-9
    # with a valid-only policy episodes end only when nothing fits (all-ones fallback mask) or on the action==A quirk
    assert s["mean_length"] > 5
    env.close()


def test_extreme_sizes(env_mod):
    """Limits of the ABI: 32x32 bins, H = 255, items up to 32 wide, bin counts that are not multiples of 32, N = 1."""
    big = [(32, 32, 200), (1, 1, 1), (7, 31, 40), (16, 16, 100), (5, 5, 5), (32, 1, 255)]
    run_against_oracle(env_mod, 33, (32, 32, 255), True, 2, dict(data_type="rs", box_set=big), T=40, seed=12)
    run_against_oracle(env_mod, 1, (10, 10, 10), False, 1, dict(data_type="rs"), T=30, seed=13)
    run_against_oracle(env_mod, 95, (20, 20, 20), True, 1, dict(data_type="rs", box_set=[(2, 2, 2), (10, 10, 3), (20, 20, 1), (9, 4, 6)]), T=40, seed=14)


@pytest.mark.parametrize("kind", ["rs", "cut2"])
def test_checkpoint_resume_is_bit_exact(env_mod, kind):
    import torch
    a = env_mod.BinPackingVecEnv(500, data_type=kind, enable_rotation=True, seed=77, want_sample=True)
    a.reset()
    for _ in range(15):
        a.step_device(a.sample_t.clone())
    sd = a.state_dict()
    ref = []
    for _ in range(20):
        a.step_device(a.sample_t.clone())
        ref.append((a.obs.clone(), a.location_masks.clone(), a.reward_t.clone(), a.done_t.clone()))
    b = env_mod.BinPackingVecEnv(500, data_type=kind, enable_rotation=True, seed=77, want_sample=True)
    b.load_state_dict(sd)
    for t in range(20):
        b.step_device(b.sample_t.clone())
        for x, y in zip(ref[t], (b.obs, b.location_masks, b.reward_t, b.done_t)):
            assert torch.equal(x, y), t
    assert a.metrics() == b.metrics()
    from b200bpp import BppError
    c = env_mod.BinPackingVecEnv(499, data_type=kind, enable_rotation=True, seed=77)
    with pytest.raises(BppError):
        c.load_state_dict(sd)


def test_single_env_cut_creators_and_registry(env_mod):
    from b200bpp import bpp0
    for kind in ("cut1", "cut2"):
        env = bpp0.make("Bpp-v0", container_size=(10, 10, 10), box_set=env_mod.default_item_set(), data_type=kind,
                        enable_rotation=False, seed=9)
        env.reset()
        seq = list(env.box_creator.box_set)
        assert sum(x * y * z for x, y, z in seq) == 1000 and all(2 <= v <= 5 for it in seq for v in it)
        want = orc.gen_episode(kind, (10, 10, 10), (2, 2, 2), (5, 5, 5), 9, 0, 0)[0]
        assert [tuple(int(v) for v in r) for r in want] == seq                 # episode 0 of global bin 0, seed 9
        assert tuple(env.box_creator.preview(3)[2]) == seq[2]
        # play a few feasible steps through the view
        for _ in range(4):
            m = env.get_possible_position().reshape(-1)
            obs, r, d, info = env.step([int(np.flatnonzero(m)[0])])
            if d:
                break
        env.box_creator.preview(len(seq) + 2)                                   # past the end: sentinel
        assert tuple(env.box_creator.box_list[-1]) == ((10, 10, 10))
        env.close()


def test_config5_full_shard_every_bin_vs_oracle(env_mod, golden_dir):
    """One whole 65 536-bin shard of BASELINE configs[4] (rank 5 of 8: rotation, cut_2 replay), every bin checked
    against the oracle at every step -- not a subsample."""
    d = np.load(os.path.join(golden_dir, "datasets.npz"))
    run_against_oracle(env_mod, 65536, (10, 10, 10), True, 2, dict(trajectories=(d["cut_2"], d["cut_2_len"]), first_traj=1),
                       T=24, seed=2024, offset=5 * 65536, global_n=524288, edge_every=6)


def test_config4_every_bin_vs_oracle(env_mod):
    """BASELINE configs[3] at full size (262 144 bins, 20x20x20, rs): every bin against the oracle for a few steps."""
    run_against_oracle(env_mod, 262144, (20, 20, 20), False, 2, dict(data_type="rs"), T=6, seed=7)


@pytest.mark.parametrize("n", [1000, 40000, 140000])
def test_step_host_pageable_buffers(env_mod, n):
    """bpp_step_host with ordinary (pageable) host memory takes the staged copy/compute pipeline (1, 2 or 4 slices);
    it must give what the device entry point gives."""
    import torch
    from b200bpp import _cabi
    a = env_mod.BinPackingVecEnv(n, enable_rotation=True, seed=3, want_sample=True)
    b = env_mod.BinPackingVecEnv(n, enable_rotation=True, seed=3, want_sample=True)
    a.reset(); b.reset()
    lib = _cabi.lib()
    P = lambda t: C.c_void_p(t.data_ptr())
    N = lambda x: x.ctypes.data_as(C.c_void_p)
    for t in range(5):
        act = a.sample_t.cpu().numpy().copy()                 # pageable
        rew, done = np.empty(n, np.float32), np.empty(n, np.uint8)
        cnt, rat = np.empty(n, np.int32), np.empty(n, np.float32)
        _cabi.check(lib.bpp_step_host(b._h, N(act), P(b.obs), P(b.location_masks), N(rew), N(done), N(cnt), N(rat),
                                      P(b.sample_t), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        a.step_device(torch.as_tensor(act, device="cuda"))
        assert torch.equal(a.obs, b.obs) and torch.equal(a.location_masks, b.location_masks) and torch.equal(a.sample_t, b.sample_t)
        assert np.array_equal(t2n(a.reward_t), rew) and np.array_equal(t2n(a.done_t), done)
        assert np.array_equal(t2n(a.counter_t), cnt) and np.array_equal(t2n(a.ratio_t), rat)
    assert a.metrics() == b.metrics()


@pytest.mark.parametrize("rot", [False, True])
@pytest.mark.parametrize("variant", [2, 1])
def test_headline_shape_every_bin_vs_oracle(env_mod, rot, variant):
    """The workload bench.py times (2^19 bins, 10x10x10, rs; kernel instantiation <ROT, VARIANT, true, 10, true> with
    32-bin warp batches): every bin against the oracle at every step, rotation off and on, mask rules v2 and v1."""
    m = run_against_oracle(env_mod, 1 << 19, (10, 10, 10), rot, variant, dict(data_type="rs"), T=6, seed=1234, edge_every=3)
    assert m["steps"] == 6 << 19


@pytest.mark.parametrize("n,rot", [(4096, False), (4096, True), (150000, False), (70001, True)])
def test_packed_results_and_multi_step_launch(env_mod, n, rot):
    """bpp_step_packed: 16-byte result records instead of four arrays, K env steps in one launch (every step after the
    first takes the action the previous one sampled).  4096 bins: every warp owns one batch and keeps its tiles in shared
    memory across the steps; 150 000 bins: 32-bin batches, tiles reloaded every step; 70 001: ragged 16-bin batches."""
    import torch
    from b200bpp import _cabi
    K = 4
    env = env_mod.BinPackingVecEnv(n, enable_rotation=rot, seed=99, want_sample=True)
    o = orc.OracleEnv(n, (10, 10, 10), rot=rot, seed=99, n_threads=0)
    env.reset(); o.reset()
    lib = _cabi.lib()
    P = lambda t: C.c_void_p(t.data_ptr())
    res = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    for rnd in range(3):
        assert np.array_equal(t2n(env.sample_t), o.sample)
        _cabi.check(lib.bpp_step_packed(env._h, P(env.sample_t), P(env.obs), P(env.location_masks), P(res), P(env.done_t),
                                        P(env.sample_t), K, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        for _ in range(K):
            o.step(o.sample.copy())
        rec = t2n(res)
        assert np.array_equal(t2n(env.obs), o.obs) and np.array_equal(t2n(env.location_masks), o.mask), rnd
        assert np.array_equal(rec[:, 0].copy().view(np.float32), o.reward) and np.array_equal(rec[:, 2], o.counter)
        assert np.array_equal(rec[:, 3].astype(np.int64), o.sample) and np.array_equal(t2n(env.done_t), o.done)
        assert np.max(np.abs(rec[:, 1] / 1000.0 - o.ratio)) <= 1e-6
    plain, box, counter, vol = env.export_state()
    oplain, obox, ocounter, ovol = o.get_state()
    assert np.array_equal(t2n(plain), oplain) and np.array_equal(t2n(box), obox) and np.array_equal(t2n(vol), ovol)
    assert env.metrics() == o.metrics()
    env.close()


def test_host_step_views_and_exact_episode_returns(env_mod):
    """BinPackingVecEnv.step(): reward / done / infos are views of NUMA-local pinned records (valid until the second call
    after); device and host actions give the same step; info['episode']['r'] is Monitor's round(sum of float64 rewards, 6)
    for a bin whose volume is not a power of ten."""
    import torch
    size = (8, 12, 9)
    a = env_mod.BinPackingVecEnv(600, size, enable_rotation=True, seed=5, want_sample=True)
    b = env_mod.BinPackingVecEnv(600, size, enable_rotation=True, seed=5, want_sample=True)
    o = orc.OracleEnv(600, size, rot=True, seed=5, n_threads=0)
    a.reset(); b.reset(); o.reset()
    ret = np.zeros(600)
    for t in range(30):
        act = o.sample.copy()
        _, ra, da, ia = a.step(torch.as_tensor(act, device="cuda").view(-1, 1))      # a GPU policy's actions
        _, rb, db, ib = b.step(act.reshape(-1, 1))                                   # host actions
        o.step(act)
        assert ra.shape == (600, 1) and not ra.is_cuda and da.dtype == np.bool_
        assert np.array_equal(ra.numpy().reshape(-1), o.reward) and np.array_equal(rb.numpy().reshape(-1), o.reward)
        assert np.array_equal(da.astype(np.uint8), o.done) and np.array_equal(db.astype(np.uint8), o.done)
        assert torch.equal(a.obs, b.obs) and torch.equal(a.location_masks, b.location_masks) and torch.equal(a.sample_t, b.sample_t)
        ret += np.where(o.done == 1, 0.0, o.reward.astype(np.float64))
        for i in np.flatnonzero(o.done):
            vol = int(round(float(o.ratio[i]) * 864))
            assert ia[int(i)]["episode"]["r"] == round(10.0 * vol / 864, 6) and ia[int(i)]["episode"]["l"] == int(o.counter[i]) + 1
            assert abs(ia[int(i)]["episode"]["r"] - ret[i]) < 5e-6                   # float32 rewards summed vs exact
            assert ib[int(i)]["ratio"] == vol / 864
            ret[i] = 0.0
    a.close(); b.close()


def test_int32_strided_actions_from_previous_records(env_mod):
    """bpp_step_packed_host_i32: int32 actions read with a stride from pinned host memory -- here the next_action column
    of the records the previous step wrote -- give the same rollout as int64 actions through bpp_step_packed_host."""
    import torch
    from b200bpp import _cabi
    from b200bpp.vec_env import _HostBuffer
    n = 3000
    a = env_mod.BinPackingVecEnv(n, enable_rotation=True, seed=8, want_sample=True)
    b = env_mod.BinPackingVecEnv(n, enable_rotation=True, seed=8, want_sample=True)
    a.reset(); b.reset()
    lib = _cabi.lib()
    P = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    hb = [_HostBuffer(lib, 0, 16 * n), _HostBuffer(lib, 0, 16 * n), _HostBuffer(lib, 0, n), _HostBuffer(lib, 0, 8 * n),
          _HostBuffer(lib, 0, 16 * n), _HostBuffer(lib, 0, n)]
    rec = [h.view(np.int32, (n, 4)) for h in hb[:2]]
    rec[0][:, 3] = t2n(a.sample_t).astype(np.int32)
    act64, rec_b, done_b = hb[3].view(np.int64, (n,)), hb[4].view(np.int32, (n, 4)), hb[5].view(np.uint8, (n,))
    for t in range(12):
        act64[:] = rec[t & 1][:, 3]
        _cabi.check(lib.bpp_step_packed_host_i32(a._h, C.c_void_p(hb[t & 1].ptr + 12), 4, P(a.obs), P(a.location_masks),
                                                 C.c_void_p(hb[(t + 1) & 1].ptr), C.c_void_p(hb[2].ptr), None, st))
        _cabi.check(lib.bpp_step_packed_host(b._h, C.c_void_p(hb[3].ptr), P(b.obs), P(b.location_masks), C.c_void_p(hb[4].ptr),
                                             C.c_void_p(hb[5].ptr), None, st))
        assert np.array_equal(rec[(t + 1) & 1], rec_b) and np.array_equal(hb[2].view(np.uint8, (n,)), done_b)
        assert torch.equal(a.obs, b.obs) and torch.equal(a.location_masks, b.location_masks)
    assert a.metrics() == b.metrics() and a.metrics()["episodes"] > 0
    a.close(); b.close()
    for h in hb:
        h.free()


def test_steps_are_cuda_graph_capturable(env_mod):
    """After one ordinary call (which sizes the launch), bpp_step makes no call that a stream capture forbids: 8 steps
    captured into one CUDA graph and replayed give what 8 plain calls give (small envs are launch-bound; a graph or
    bpp_step_packed's n_steps remove the per-step launch cost)."""
    import torch
    a = env_mod.BinPackingVecEnv(4096, enable_rotation=True, seed=11, want_sample=True)
    b = env_mod.BinPackingVecEnv(4096, enable_rotation=True, seed=11, want_sample=True)
    a.reset(); b.reset()
    for env in (a, b):
        env.step_device(env.sample_t)               # in place: the kernel reads a bin's action before it writes its next one
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(8):
                a.step_device(a.sample_t)
    for rnd in range(3):
        g.replay()
        for _ in range(8):
            b.step_device(b.sample_t)
        torch.cuda.synchronize()
        assert torch.equal(a.obs, b.obs) and torch.equal(a.location_masks, b.location_masks) and torch.equal(a.sample_t, b.sample_t)
        assert torch.equal(a.reward_t, b.reward_t) and torch.equal(a.done_t, b.done_t)
    ma, mb = a.metrics(), b.metrics()
    assert (ma["episodes"], ma["sum_counter"], ma["sum_vol"]) == (mb["episodes"], mb["sum_counter"], mb["sum_vol"])
    a.close(); b.close()


def test_packed_host_step_with_ordinary_memory(env_mod):
    """bpp_step_packed_host with pageable host buffers (numpy arrays): staged through device copies, same records as the
    zero-copy route; the int32 entry point refuses ordinary memory loudly."""
    import torch
    from b200bpp import _cabi, BppError
    n = 5000
    a = env_mod.BinPackingVecEnv(n, seed=17, want_sample=True)
    b = env_mod.BinPackingVecEnv(n, seed=17, want_sample=True)
    a.reset(); b.reset()
    lib = _cabi.lib()
    P = lambda t: C.c_void_p(t.data_ptr())
    Np = lambda x: x.ctypes.data_as(C.c_void_p)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for t in range(6):
        act = t2n(a.sample_t).copy()
        rec, done = np.zeros((n, 4), np.int32), np.zeros(n, np.uint8)
        _cabi.check(lib.bpp_step_packed_host(b._h, Np(act), P(b.obs), P(b.location_masks), Np(rec), Np(done), P(b.sample_t), st))
        _, reward, d, infos = a.step(torch.as_tensor(act, device="cuda").view(-1, 1))
        assert np.array_equal(rec[:, 0].copy().view(np.float32), reward.numpy().reshape(-1)) and np.array_equal(done, d.astype(np.uint8))
        assert np.array_equal(rec[:, 3].astype(np.int64), t2n(a.sample_t)) and torch.equal(a.obs, b.obs)
        assert all(infos[i]["counter"] == int(rec[i, 2]) for i in range(0, n, 997))
    with pytest.raises(BppError):
        act32 = np.zeros(n, np.int32)
        _cabi.check(lib.bpp_step_packed_host_i32(b._h, Np(act32), 1, P(b.obs), P(b.location_masks), Np(rec), Np(done), None, st))
    a.close(); b.close()
