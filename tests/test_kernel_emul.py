"""The CUDA kernel SOURCES, interpreted lane by lane on the CPU (tests/simt/), against the oracle and the reference's
golden traces.

This is not a CPU path of the product (libbpp_b200.so has none): g++ compiles csrc/*.cu against a SIMT shim so that the
no-GPU suite exercises the real kernels -- template dispatch, 32-bin warp batches, shared-memory staging, warp
collectives, the generators' draw order -- bit for bit before a B200 is spent on them.  The `-m gpu` tests remain the
parity tests proper; sizes here are small because every warp collective costs 32 fiber switches.
"""
import os
import sys

import numpy as np
import pytest

import oracle as orc

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt"))
import emul   # noqa: E402


# kernel builds (switches in csrc/bpp_common.cuh, online-3d-bpp-drl_b200/build.py -D...): the default one (packed
# 16-bit-pair mask sweep, lane-per-bin placement check, single sweep for square items under rotation) and the one with
# every switch off, which tools/ab_bench.py keeps as the timing baseline
BASELINE = ("BPP_PACKED_SWEEP=0", "BPP_LANE_CHECK=0", "BPP_SQUARE_COPY=0", "BPP_PLAIN_STORES=0")
VARIANTS = [pytest.param((), id="default"), pytest.param(BASELINE, id="baseline")]


def run_against_oracle(N, size, rot, variant, gen_kw, T, seed, offset=0, global_n=None, edge_every=0, num_sms=2,
                       blocks_per_sm=1, defines=()):
    """Same protocol as tests/test_gpu_parity.py::run_against_oracle, with the interpreted kernels in place of the GPU."""
    e = emul.EmulEnv(N, size, rot=rot, mask_variant=variant, seed=seed, global_offset=offset, global_n=global_n,
                     num_sms=num_sms, blocks_per_sm=blocks_per_sm, defines=defines, **gen_kw)
    o = orc.OracleEnv(N, size, rot=rot, mask_variant=variant, seed=seed, global_offset=offset, global_n=global_n,
                      n_threads=1, **gen_kw)
    e.reset()
    o.reset()
    A = size[0] * size[1]
    rng = np.random.default_rng(seed)
    for t in range(T + 1):
        assert np.array_equal(e.obs, o.obs), t
        assert np.array_equal(e.mask, o.mask), t
        assert np.array_equal(e.sample, o.sample), t
        if t == T:
            break
        actions = o.sample.copy()
        if edge_every and t % edge_every == 0:      # out-of-range / quirk actions (action == A, negative, past the end)
            k = rng.integers(0, N, max(1, N // 16))
            actions[k] = rng.choice([A, -1, A * (2 if rot else 1) + 3, 0] if rot else [A, -1, 0], size=k.shape[0])
        e.step(actions)
        o.step(actions)
        assert np.array_equal(e.done, o.done), t
        assert np.array_equal(e.counter, o.counter), t
        assert np.array_equal(e.reward, o.reward), t
        assert np.max(np.abs(e.ratio.astype(np.float64) - o.ratio)) <= 1e-6
    for a, b in zip(e.get_state(), o.get_state()):
        assert np.array_equal(a, b)
    assert e.metrics_dict() == o.metrics()
    return e.metrics_dict()


@pytest.mark.parametrize("defines", VARIANTS)
@pytest.mark.parametrize("rot,variant", [(False, 2), (True, 2), (True, 1)])
def test_step_kernel_rs_10cube(rot, variant, defines):
    # the compile-time 10x10 fast path (CS = 10, FAST), 32-bin warp batches with a ragged tail, several batches per warp
    m = run_against_oracle(300, (10, 10, 10), rot, variant, dict(gen="rs"), T=24, seed=1234, edge_every=7, offset=77,
                           global_n=4096, defines=defines)
    assert m["episodes"] > 300


@pytest.mark.parametrize("defines", VARIANTS)
def test_step_kernel_20cube_and_small_batches(defines):
    # CS = 20; few bins on "many SMs" -> the batch-size rule picks fewer bins per warp
    run_against_oracle(40, (20, 20, 20), False, 2, dict(gen="rs"), T=40, seed=99, num_sms=8, blocks_per_sm=2, defines=defines)


@pytest.mark.parametrize("defines", VARIANTS)
def test_step_kernel_generic_sizes(defines):
    # CS = 0: scalar stores (L*W % 4 != 0), non-square bin, items larger than the bin, footprints outside 2..5
    items = np.array([(x, y, z) for x in (1, 3, 8) for y in (2, 7) for z in (1, 4)], dtype=np.uint8)
    run_against_oracle(70, (7, 9, 6), True, 1, dict(gen="rs", item_set=items), T=30, seed=3, edge_every=4, defines=defines)
    run_against_oracle(33, (8, 12, 9), False, 2, dict(gen="rs"), T=30, seed=4, defines=defines)      # vec4 generic path


@pytest.mark.parametrize("defines", VARIANTS)
def test_step_kernel_tall_bins_and_items_that_never_fit(defines):
    # heights up to 255 (16-bit halves of the packed sweep keep their head room), items taller than the bin, footprints
    # with a side of 1 or > 5 next to the specialised 2..5 ones, rows of more than 5 cells in the lane-per-bin check
    items = np.array([(x, y, z) for x in (2, 3, 5) for y in (2, 4, 5) for z in (1, 37, 90, 200)], dtype=np.uint8)
    run_against_oracle(64, (10, 10, 255), True, 2, dict(gen="rs", item_set=items), T=30, seed=5, defines=defines)
    items = np.array([(x, y, z) for x in (1, 2, 4, 6) for y in (1, 3, 5, 6) for z in (1, 3, 6, 8)], dtype=np.uint8)
    run_against_oracle(64, (10, 10, 5), True, 2, dict(gen="rs", item_set=items), T=30, seed=6, edge_every=4, defines=defines)
    items = np.array([(x, y, z) for x in (2, 5) for y in (3, 4) for z in (1, 60, 99)], dtype=np.uint8)
    run_against_oracle(24, (20, 20, 200), True, 2, dict(gen="rs", item_set=items), T=30, seed=7, defines=defines)


@pytest.mark.parametrize("defines", VARIANTS)
def test_step_kernel_replay_rotation(golden_dir, defines):
    d = np.load(os.path.join(golden_dir, "datasets.npz"))
    run_against_oracle(96, (10, 10, 10), True, 2, dict(gen="replay", traj=d["cut_2"], traj_len=d["cut_2_len"], first_traj=1),
                       T=40, seed=8, offset=3 * 65536, global_n=524288, edge_every=5, defines=defines)


@pytest.mark.parametrize("kind,rot", [("cut1", True), ("cut2", False)])
def test_cut_generators(kind, rot):
    # bpp_gen.cu: episodes cut by one warp each, Philox draws in the oracle's order; refill list from the step kernel
    run_against_oracle(40, (10, 10, 10), rot, 2, dict(gen=kind), T=30, seed=31, offset=5000, global_n=100000, edge_every=9)


@pytest.mark.parametrize("defines", VARIANTS)
def test_golden_rollout_through_the_interpreted_kernel(golden_dir, defines):
    """The reference's own trace (recorded from the Python reference) replayed through the kernel source."""
    g = np.load(os.path.join(golden_dir, "ref_rollouts_10_rot.npz"))
    size, rot = tuple(int(v) for v in g["size"]), bool(g["rot"])
    T, N = g["action"].shape
    T = min(T, 40)
    A = size[0] * size[1]
    for variant in (1, 2):
        e = emul.EmulEnv(N, size, rot=rot, mask_variant=variant, gen="replay", traj=g["traj"], traj_len=g["traj_len"],
                         first_traj=0, defines=defines)
        want = g["mask_v1"] if variant == 1 else g["mask_v2"]
        e.reset()
        for t in range(T):
            assert np.array_equal(e.obs[:, :A], g["hm"][t].astype(np.float32)), t
            assert np.array_equal(e.mask, want[t].astype(np.float32)), t
            e.step(g["action"][t])
            assert np.array_equal(e.done, g["done"][t]) and np.array_equal(e.counter, g["counter"][t])
            assert np.array_equal(e.reward, g["reward"][t].astype(np.float32))
            assert np.max(np.abs(e.ratio.astype(np.float64) - g["ratio"][t])) <= 1e-6


@pytest.mark.parametrize("defines", VARIANTS)
def test_partial_outputs_and_observe(defines):
    """Null output pointers (the non-FAST instantiation) and bpp_observe leave the state and the other outputs intact."""
    e = emul.EmulEnv(64, (10, 10, 10), seed=5, defines=defines)
    o = orc.OracleEnv(64, (10, 10, 10), seed=5)
    e.reset(); o.reset()
    for t in range(6):
        a = o.sample.copy()
        e.step(a, want_obs=(t % 2 == 0), want_mask=(t % 3 != 0), want_sample=True, scalars=(t % 2 == 1))
        o.step(a)
        assert np.array_equal(e.sample, o.sample)
    e.observe()
    assert np.array_equal(e.obs, o.obs) and np.array_equal(e.mask, o.mask)


@pytest.mark.parametrize("size", [(10, 10, 10), (8, 12, 9)])
def test_stateless_kernels_fuzz(size):
    """bpp_mask_kernel / bpp_check_box_kernel on blocky random heightmaps, both rules, incl. the v1 != v2 region."""
    rng = np.random.default_rng(11)
    L, W, H = size
    n = 48
    plain = np.zeros((n, L, W), dtype=np.int32)
    for b in range(n):
        for _ in range(rng.integers(0, 8)):
            x, y = rng.integers(1, L + 1), rng.integers(1, W + 1)
            i, j = rng.integers(0, L - x + 1), rng.integers(0, W - y + 1)
            plain[b, i:i + x, j:j + y] = plain[b, i:i + x, j:j + y].max() + rng.integers(1, 4)
    plain = np.minimum(plain, H).reshape(n, -1)
    box = np.stack([rng.integers(1, L + 1, n), rng.integers(1, W + 1, n), rng.integers(1, 5, n)], 1).astype(np.int32)
    for rot in (False, True):
        for variant in (1, 2):
            want = np.stack([orc.mask(plain[b], box[b], size, rot, variant) for b in range(n)]).astype(np.float32)
            assert np.array_equal(emul.mask_from_plain(plain, box, size, rot, variant), want)
            obs = np.concatenate([plain] + [np.repeat(box[:, c:c + 1], L * W, 1) for c in range(3)], 1).astype(np.float32)
            assert np.array_equal(emul.mask_from_obs(obs, size, rot, variant), want)
    q = np.stack([box[:, 0], box[:, 1], rng.integers(-1, L, n), rng.integers(-1, W, n), box[:, 2]], 1).astype(np.int32)
    for variant in (1, 2):
        want = np.array([orc.check_box(plain[b], q[b, 0], q[b, 1], q[b, 2], q[b, 3], q[b, 4], size, variant) for b in range(n)])
        assert np.array_equal(emul.check_box(plain, q, size, variant), want)


@pytest.mark.parametrize("seed", range(6))
def test_random_small_configurations(seed):
    """Seeded random bin sizes / item sets / rules / rotation / generators through the generic (CS = 0) instantiations and,
    when the draw lands on them, the compile-time ones; ragged N, several launches' worth of batches."""
    rng = np.random.default_rng(1000 + seed)
    L, W = (int(v) for v in rng.integers(3, 13, 2))
    if seed % 3 == 0:
        L = W = 10                                   # the compile-time path with odd item sets
    H = int(rng.integers(4, 40))
    n_items = int(rng.integers(1, 9))
    items = np.stack([rng.integers(1, max(2, L // 2 + 2), n_items), rng.integers(1, max(2, W // 2 + 2), n_items),
                      rng.integers(1, max(2, H // 2 + 1), n_items)], 1).astype(np.uint8)
    rot, variant = bool(rng.integers(0, 2)), int(rng.integers(1, 3))
    N = int(rng.integers(1, 90))
    run_against_oracle(N, (L, W, H), rot, variant, dict(gen="rs", item_set=items), T=12, seed=seed, edge_every=5,
                       num_sms=int(rng.integers(1, 4)), blocks_per_sm=int(rng.integers(1, 3)))


def test_extreme_sizes():
    """The largest bin the library accepts (32x32x255), single-bin and 33-bin envs (one full warp batch + one bin)."""
    items = np.array([(1, 1, 1), (32, 32, 7), (5, 31, 200), (17, 2, 99), (3, 3, 255)], dtype=np.uint8)
    run_against_oracle(1, (32, 32, 255), True, 2, dict(gen="rs", item_set=items), T=12, seed=1)
    run_against_oracle(33, (32, 32, 255), False, 1, dict(gen="rs", item_set=items), T=8, seed=2, edge_every=3)
    run_against_oracle(33, (1, 1, 1), True, 2, dict(gen="rs", item_set=np.array([(1, 1, 1)], dtype=np.uint8)), T=6, seed=3)


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("defines", VARIANTS)
def test_lane_schedule_does_not_matter(order, defines):
    """Between two warp collectives the interpreter runs the lanes one after another; a shared-memory write that another
    lane reads without a __syncwarp() in between would depend on that order.  Same results with the lanes reversed and
    with the odd lanes first (step kernel with rotation, both generators, stateless masks)."""
    emul.set_lane_order(order, defines)
    try:
        run_against_oracle(100, (10, 10, 10), True, 2, dict(gen="rs"), T=16, seed=21, edge_every=5, defines=defines)
        run_against_oracle(40, (20, 20, 20), False, 1, dict(gen="rs"), T=10, seed=22, num_sms=4, defines=defines)
        run_against_oracle(40, (7, 9, 6), True, 2, dict(gen="rs", item_set=np.array([(1, 2, 1), (3, 3, 2), (7, 2, 1)], np.uint8)),
                           T=10, seed=23, defines=defines)
        if not defines:
            run_against_oracle(24, (10, 10, 10), True, 2, dict(gen="cut1"), T=12, seed=24)
            run_against_oracle(24, (10, 10, 10), False, 2, dict(gen="cut2"), T=12, seed=25)
    finally:
        emul.set_lane_order(0, defines)


def test_rule_v1_instantiations():
    """Rule v1 (Space.check_box as mask rule) on the compile-time sizes and on a size without 128-bit stores: template
    instantiations the other cases leave out (gcov over the interpreted build: all 24 step-kernel instantiations run)."""
    run_against_oracle(96, (10, 10, 10), False, 1, dict(gen="rs"), T=16, seed=41, edge_every=5)
    run_against_oracle(40, (20, 20, 20), True, 1, dict(gen="rs"), T=12, seed=42, num_sms=4)
    items = np.array([(1, 2, 1), (3, 3, 2), (6, 2, 1), (2, 8, 3)], dtype=np.uint8)
    run_against_oracle(50, (7, 9, 6), False, 1, dict(gen="rs", item_set=items), T=16, seed=43, edge_every=4)


def test_philox_streams():
    """The hand-written Philox4x32-10 of the kernels against the oracle's independent C copy (which tests/test_oracle_golden.py
    pins to the Random123 known answers); the GPU suite adds cuRAND's own generator."""
    for seed, dom, aux, gid in [(1234, 0, 0, 0), (1234, 0, 0, 77), (0xDEADBEEFCAFE, 1, 0, 2 ** 33 + 5), (7, 2, 12345, 9)]:
        got = emul.debug_draws(seed, dom, aux, gid, 70)
        want = np.array([orc.draw(seed, dom, aux, gid, i) for i in range(70)], dtype=np.uint32)
        assert np.array_equal(got, want)


NO_TMA = ("BPP_TMA=0",)


@pytest.mark.parametrize("defines", [pytest.param((), id="tma"), pytest.param(NO_TMA, id="no-tma")])
@pytest.mark.parametrize("rot", [False, True])
def test_packed_results_and_multi_step_launch(rot, defines):
    """bpp_result records (reward | placed volume | counter | next action) instead of four arrays, and K env steps in ONE
    launch (policy-free rollouts: every step after the first takes the action the previous step sampled) against K single
    steps of the oracle -- on the bulk-copy path (10x10, every output) and with that path compiled out."""
    size, K = (10, 10, 10), 5
    # 150 bins on 16 resident warps: several batches per warp, tiles reloaded every step; 100 bins on 128 warps: every
    # warp owns at most one batch and keeps its tiles in shared memory across the K steps
    for N, num_sms, blocks, gen_kw in ((150, 2, 1, dict(gen="rs")), (100, 8, 2, dict(gen="rs"))):
        e = emul.EmulEnv(N, size, rot=rot, seed=21, defines=defines, num_sms=num_sms, blocks_per_sm=blocks, **gen_kw)
        o = orc.OracleEnv(N, size, rot=rot, seed=21, n_threads=1, **gen_kw)
        e.reset()
        o.reset()
        for rnd in range(3):
            a0 = o.sample.copy()
            e._call(0, np.ascontiguousarray(a0), packed=True, n_steps=K)
            for _ in range(K):
                o.step(o.sample.copy())
            assert np.array_equal(e.obs, o.obs) and np.array_equal(e.mask, o.mask), rnd
            rec = e.packed
            assert np.array_equal(rec[:, 0].view(np.float32), o.reward)
            assert np.array_equal(rec[:, 2], o.counter) and np.array_equal(rec[:, 3].astype(np.int64), o.sample)
            assert np.array_equal(e.done, o.done)
            vol = rec[:, 1].astype(np.float64) / 1000.0
            assert np.max(np.abs(vol - o.ratio)) <= 1e-6
            assert np.array_equal(e.sample, o.sample)          # `sample` is still written next to the records
        # int32 actions with a stride: the next_action column of the last records drives one more step
        prev = e.packed.copy()
        e._call(0, None, packed=True, actions32=prev[:, 3], actions32_stride=4)
        o.step(o.sample.copy())
        assert np.array_equal(e.obs, o.obs) and np.array_equal(e.mask, o.mask)
        assert np.array_equal(e.packed[:, 0].view(np.float32), o.reward) and np.array_equal(e.done, o.done)
        for a, b in zip(e.get_state(), o.get_state()):
            assert np.array_equal(a, b)
        m = e.metrics_dict()
        m["steps"] = o.metrics()["steps"]                      # the emulated env counts launches, not steps
        assert m == o.metrics()


def test_bulk_copy_path_32_bin_batches():
    """Envs big enough to fill the SMs with 32-bin batches take the <.., TMA = 5> instantiation (3 CTAs per SM); with
    2 emulated SMs that is 1216 bins.  Ragged last batch, several batches per warp, both mask rules."""
    for variant in (2, 1):
        run_against_oracle(1300, (10, 10, 10), False, variant, dict(gen="rs"), T=5, seed=77, edge_every=2, num_sms=2,
                           blocks_per_sm=3)


def test_bulk_copy_path_cta_sizes():
    """The 10x10 bulk-copy instantiations differ in CTA size (launch_rv): 7 warps with rotation once 16-bin batches fill
    the SMs (1 emulated SM: 28 batches = 433+ bins), 6 warps below that down to 160 bins per SM, 8 warps for smaller envs
    and for the 32-bin batches above.  Ragged last batch, several batches per warp, fewer CTAs than the kernel was
    bounded for."""
    run_against_oracle(470, (10, 10, 10), True, 2, dict(gen="rs"), T=5, seed=78, edge_every=2, num_sms=1, blocks_per_sm=2)
    run_against_oracle(200, (10, 10, 10), True, 1, dict(gen="rs"), T=5, seed=79, edge_every=2, num_sms=1, blocks_per_sm=2)
    run_against_oracle(200, (10, 10, 10), False, 2, dict(gen="cut2"), T=6, seed=80, num_sms=1, blocks_per_sm=3)
    run_against_oracle(150, (10, 10, 10), True, 2, dict(gen="rs"), T=5, seed=81, edge_every=2, num_sms=1, blocks_per_sm=2)
    # the remaining (rotation, rule, CTA size) combinations, so that every instantiation of the default build runs
    run_against_oracle(200, (10, 10, 10), False, 1, dict(gen="rs"), T=4, seed=82, edge_every=2, num_sms=1, blocks_per_sm=2)
    run_against_oracle(470, (10, 10, 10), True, 1, dict(gen="rs"), T=4, seed=83, edge_every=2, num_sms=1, blocks_per_sm=2)
    run_against_oracle(200, (10, 10, 10), True, 2, dict(gen="cut1"), T=4, seed=84, num_sms=1, blocks_per_sm=3)
