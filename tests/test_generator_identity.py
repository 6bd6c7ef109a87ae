"""cut1 / cut2 episode generators: IDENTICAL episodes to the unmodified reference generators on identical random draws.

tests/golden/gen_identity.npz holds episodes produced by the reference's own `CuttingBoxCreator` (cutCreator.py:32-128)
and `MDlayerBoxCreator` (mdCreator.py:59-166) with `random.choice` / `random.randint` / `np.random.rand` fed from the
Philox generator stream of (seed, global bin id, episode) -- see tests/golden/make_gen_identity.py.  The oracle, the CUDA
kernel sources under CPU interpretation and (GPU test) the device generator must reproduce them item for item: draw
order of the level-by-level split sweeps, axis choice, cut positions, candidate pops, rotation coins, cut2's
remove-while-iterating traversal and the tie order of its stable sort.
"""
import os
import sys

import numpy as np
import pytest

import oracle as orc

LO, HI = (2, 2, 2), (5, 5, 5)


def load_cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "gen_identity.npz"))
    out = []
    for ci, row in enumerate(g["cases"]):
        kind, size, rot, seed, gid0, n_bins, n_ep = row
        out.append(dict(kind=str(kind), size=tuple(int(v) for v in str(size).split("x")), rot=bool(int(rot)), seed=int(seed),
                        gid0=int(gid0), n_bins=int(n_bins), n_ep=int(n_ep), items=g["items_%d" % ci], lens=g["len_%d" % ci]))
    return out


def ref_episode(case, b, e):
    i = b * case["n_ep"] + e
    return case["items"][i, :case["lens"][i]]


def test_oracle_equals_reference_episodes(golden_dir):
    n = 0
    for c in load_cases(golden_dir):
        for b in range(c["n_bins"]):
            for e in range(c["n_ep"]):
                items, _ = orc.gen_episode(c["kind"], c["size"], LO, HI, c["seed"], c["gid0"] + b, e, rot=c["rot"])
                assert np.array_equal(items, ref_episode(c, b, e)), (c["kind"], c["size"], b, e)
                n += 1
    assert n >= 3000


def test_interpreted_kernel_equals_reference_episodes(golden_dir):
    """csrc/bpp_gen.cu, compiled for the CPU (tests/simt): buffers 0 / 1 after reset() hold episodes 0 / 1 of every bin."""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "simt"))
    import emul
    for c in load_cases(golden_dir):
        nb = min(c["n_bins"], 48 if c["size"][0] <= 12 else 6)
        env = emul.EmulEnv(nb, c["size"], rot=c["rot"], gen=c["kind"], seed=c["seed"], global_offset=c["gid0"], global_n=1 << 40)
        env.reset()
        for b in range(nb):
            for e in range(2):
                n = int(env.ep_len2[e, b])
                w = env.ep_items[e, b, :n]
                got = np.stack([w & 0xff, (w >> 8) & 0xff, (w >> 16) & 0xff], 1).astype(np.uint8)
                assert np.array_equal(got, ref_episode(c, b, e)), (c["kind"], c["size"], b, e)


def test_live_reference_identity_on_fresh_streams():
    """Where the reference tree is present: new seeds / bins / episodes through the patched reference vs the oracle."""
    import reference_loader as rl
    ref = rl.load_reference()
    if ref is None:
        pytest.skip("reference tree not present (GPU box): the committed fixture covers this")
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_gen_identity as mg
    stream = mg.PhiloxStream()
    with mg.patched(stream):
        for kind, size, rot in (("cut1", (10, 10, 10), True), ("cut2", (10, 10, 10), False), ("cut1", (6, 7, 11), False),
                                ("cut2", (12, 8, 10), False)):
            stream.start(1, 0, 0)
            with rl.quiet():
                creator = ref.CuttingBoxCreator(list(size), LO + HI, rot) if kind == "cut1" else ref.MDlayerBoxCreator(list(size), (2, 5))
            for trial in range(40):
                seed, gid, ep = 1000 + trial, (trial * 7919) << (trial % 30), trial % 5
                want = mg.reference_episode(ref, creator, kind, size, stream, seed, gid, ep)
                got, _ = orc.gen_episode(kind, size, LO, HI, seed, gid, ep, rot=rot)
                assert [tuple(int(v) for v in r) for r in got] == want, (kind, size, trial)


@pytest.mark.gpu
def test_device_generators_equal_reference_episodes(golden_dir):
    """The device generator (bpp_gen_kernel) against the reference-produced episodes, every bin of every case."""
    import b200bpp
    for c in load_cases(golden_dir):
        env = b200bpp.BinPackingVecEnv(c["n_bins"], c["size"], enable_rotation=c["rot"], data_type=c["kind"], seed=c["seed"],
                                       global_offset=c["gid0"], global_n=1 << 40)
        for e in range(c["n_ep"]):
            env.reset()                         # first reset: episode 0; an explicit second reset moves on to episode 1
            for b in range(c["n_bins"]):
                got = np.asarray(env.episode_items(b), np.uint8).reshape(-1, 3)
                assert np.array_equal(got, ref_episode(c, b, e)), (c["kind"], c["size"], b, e)
        env.close()

