#!/usr/bin/env python
"""Episodes of the UNMODIFIED reference generators driven by the oracle's Philox stream -> tests/golden/gen_identity.npz.

    python tests/golden/make_gen_identity.py          # build container only (needs /root/reference)

The reference's cut generators (envs/bpp0/cutCreator.py:32-128 `CuttingBoxCreator`, envs/bpp0/mdCreator.py:59-166
`MDlayerBoxCreator`) take their randomness from `random.choice`, `random.randint` and `np.random.rand`.  Here those
three entry points are replaced -- in the interpreter, the reference files stay untouched -- by functions that consume
the generator stream of the device / oracle implementation (Philox4x32-10, domain 2, aux = episode, subsequence = global
bin id; DESIGN.md "RNG streams") with the same reduction the kernels use:

    random.choice(seq)   -> seq[mulhi32(u, len(seq))]
    random.randint(a, b) -> a + mulhi32(u, b - a + 1)
    np.random.rand()     -> u / 2^32            (rd < 0.5  <=>  u < 2^31)

With that, the reference produces -- draw for draw -- the episodes the oracle (oracle/bpp_oracle.c orc_gen_cut1/2) and
the device generator (csrc/bpp_gen.cu) must produce.  The recorded episodes travel to the GPU box as a fixture:
tests/test_generator_identity.py compares oracle and device against them.
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import reference_loader as rl   # noqa: E402
import oracle as orc            # noqa: E402

DOM_GEN = 2
LO, HI = (2, 2, 2), (5, 5, 5)
# (kind, container, rotation, seed, first global bin id, number of bins, episodes per bin)
CASES = [
    ("cut1", (10, 10, 10), False, 11, 0, 600, 2),
    ("cut1", (10, 10, 10), True, 12, 70000, 300, 2),
    ("cut2", (10, 10, 10), False, 13, 0, 600, 2),
    ("cut1", (20, 20, 20), False, 14, 5, 40, 2),
    ("cut2", (20, 20, 20), False, 15, 5, 40, 2),
    ("cut1", (8, 12, 9), True, 16, 1 << 33, 60, 2),
    ("cut2", (8, 12, 9), False, 17, 1 << 33, 60, 2),
]


class PhiloxStream(object):
    """The (seed, gid, episode) generator stream, one 32-bit draw per call."""

    def __init__(self):
        self.seed = self.gid = self.episode = 0
        self.k = 0

    def start(self, seed, gid, episode):
        self.seed, self.gid, self.episode, self.k = seed, gid, episode, 0

    def u32(self):
        v = orc.draw(self.seed, DOM_GEN, self.episode, self.gid, self.k)
        self.k += 1
        return v


def patched(stream):
    """Context manager: random.choice / random.randint / np.random.rand consume `stream`."""
    import contextlib

    @contextlib.contextmanager
    def cm():
        saved = (random.choice, random.randint, np.random.rand)
        random.choice = lambda seq: seq[(stream.u32() * len(seq)) >> 32]
        random.randint = lambda a, b: a + ((stream.u32() * (b - a + 1)) >> 32)
        np.random.rand = lambda *a: stream.u32() / 4294967296.0
        try:
            yield
        finally:
            random.choice, random.randint, np.random.rand = saved
    return cm()


def reference_episode(ref, creator, kind, size, stream, seed, gid, episode):
    """One episode of the reference creator on the (seed, gid, episode) stream: list of (x, y, z)."""
    stream.start(seed, gid, episode)
    with rl.quiet():
        creator.reset()
    items = []
    if kind == "cut1":
        # generate_box_size pops one ready piece per call (cutCreator.py:112-128); with no candidate left it emits
        # the bin-size sentinel from then on (:113-115)
        while len(creator.candidates) > 0:
            creator.generate_box_size()
            items.append(tuple(int(v) for v in creator.box_list[-1]))
        assert len(creator.meta_list) == 0 and sum(x * y * z for x, y, z in items) == size[0] * size[1] * size[2]
        creator.generate_box_size()
        assert tuple(creator.box_list[-1]) == tuple(size)
    else:
        # MDlayerBoxCreator.reset() leaves the sorted episode + [10,10,10] in box_set (mdCreator.py:153-162)
        items = [tuple(int(v) for v in b) for b in creator.box_set[:-1]]
        assert list(creator.box_set[-1]) == [10, 10, 10]
    return items


def main():
    ref = rl.load_reference()
    if ref is None:
        sys.exit("reference tree not found; this fixture can only be regenerated in the build container")
    stream = PhiloxStream()
    out = {"cases": np.array([[c[0], "x".join(map(str, c[1])), int(c[2]), c[3], c[4], c[5], c[6]] for c in CASES], dtype=object).astype(str)}
    with patched(stream):
        for ci, (kind, size, rot, seed, gid0, n_bins, n_ep) in enumerate(CASES):
            stream.start(seed, 0, 0xFFFFFFF)       # the constructors also cut one episode: give them a throw-away stream
            with rl.quiet():
                if kind == "cut1":
                    creator = ref.CuttingBoxCreator(list(size), LO + HI, rot)
                else:
                    creator = ref.MDlayerBoxCreator(list(size), (LO[0], HI[0]))
            eps = []
            for b in range(n_bins):
                for e in range(n_ep):
                    eps.append(reference_episode(ref, creator, kind, size, stream, seed, gid0 + b, e))
            kmax = max(len(e) for e in eps)
            arr = np.zeros((len(eps), kmax, 3), np.uint8)
            lens = np.zeros(len(eps), np.int32)
            for i, e in enumerate(eps):
                arr[i, :len(e)] = np.asarray(e, np.uint8).reshape(-1, 3)
                lens[i] = len(e)
            out["items_%d" % ci], out["len_%d" % ci] = arr, lens
            print("%s %s rot=%d: %d episodes, %d..%d items" % (kind, size, rot, len(eps), lens.min(), lens.max()))
    np.savez_compressed(os.path.join(HERE, "gen_identity.npz"), **out)


if __name__ == "__main__":
    main()
