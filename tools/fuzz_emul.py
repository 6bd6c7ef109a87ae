#!/usr/bin/env python
"""Randomised campaign: the interpreted kernels (tests/simt) against the oracle over random configurations.

    python tools/fuzz_emul.py [--minutes 10] [--seed 0] [--defines BPP_TMA=0 ...]

Bin sizes 1..32, heights 1..255, item sets with dims 1..(side+2), both mask rules, rotation, rs / replay / cut1 / cut2,
ragged N, random grid shapes, out-of-range actions, and the three lane schedules of the interpreter.  Stops at the first
mismatch with the configuration that produced it.  Not part of the test suite (it is as long as you let it run).
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "simt"), os.path.join(ROOT, "oracle"),
          os.path.join(ROOT, "online-3d-bpp-drl_b200")):
    sys.path.insert(0, p)

import numpy as np          # noqa: E402
import emul                 # noqa: E402
import test_kernel_emul as T   # noqa: E402


def one(rng, defines):
    kind = rng.choice(["rs", "rs", "rs", "replay", "cut1", "cut2"])
    if kind in ("cut1", "cut2"):
        side = int(rng.choice([6, 8, 10, 12]))
        L = W = side
        H = int(rng.choice([6, 8, 10, 12]))
        lo = int(rng.integers(1, 4))
        hi = int(rng.integers(2 * lo - 1 if lo > 1 else 1, 2 * lo + 3))
        hi = max(hi, 2 * lo - 1, lo)
        gen_kw = dict(gen=kind, lo=(lo, lo, lo), hi=(hi, hi, hi))
    else:
        L, W = (int(v) for v in rng.integers(1, 33, 2))
        if rng.random() < 0.4:
            L = W = int(rng.choice([10, 20]))
        H = int(rng.integers(1, 256)) if rng.random() < 0.3 else int(rng.integers(1, 25))
        n_items = int(rng.integers(1, 10))
        items = np.stack([rng.integers(1, min(L + 2, 33), n_items), rng.integers(1, min(W + 2, 33), n_items),
                          rng.integers(1, min(H + 2, 256), n_items)], 1).astype(np.uint8)
        if kind == "rs":
            gen_kw = dict(gen="rs", item_set=items)
        else:
            n_traj, cap = int(rng.integers(1, 7)), int(rng.integers(1, 9))
            traj = items[rng.integers(0, n_items, (n_traj, cap))]
            gen_kw = dict(gen="replay", traj=traj, traj_len=rng.integers(0, cap + 1, n_traj).astype(np.int32),
                          first_traj=int(rng.integers(0, 5)))
    cfg = dict(N=int(rng.integers(1, 120)), size=(L, W, H), rot=bool(rng.integers(0, 2)), variant=int(rng.integers(1, 3)),
               gen_kw=gen_kw, T=int(rng.integers(2, 16)), seed=int(rng.integers(0, 1 << 30)), offset=int(rng.integers(0, 1 << 20)),
               edge_every=int(rng.integers(0, 6)), num_sms=int(rng.integers(1, 5)), blocks_per_sm=int(rng.integers(1, 4)))
    order = int(rng.integers(0, 3))
    emul.set_lane_order(order, defines)
    try:
        T.run_against_oracle(cfg["N"], cfg["size"], cfg["rot"], cfg["variant"], cfg["gen_kw"], cfg["T"], cfg["seed"],
                             offset=cfg["offset"], global_n=cfg["offset"] + cfg["N"] + int(rng.integers(0, 1000)),
                             edge_every=cfg["edge_every"], num_sms=cfg["num_sms"], blocks_per_sm=cfg["blocks_per_sm"],
                             defines=defines)
    except Exception:
        print("MISMATCH / ERROR with lane order %d and configuration:\n%r" % (order, cfg), flush=True)
        raise
    finally:
        emul.set_lane_order(0, defines)
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--minutes", type=float, default=10.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--defines", nargs="*", default=[])
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    t0, n, steps = time.time(), 0, 0
    while time.time() - t0 < 60 * a.minutes:
        cfg = one(rng, tuple(a.defines))
        n += 1
        steps += cfg["N"] * cfg["T"]
        if n % 25 == 0:
            print("%d configurations, %d env-steps, %.0f s" % (n, steps, time.time() - t0), flush=True)
    print("done: %d configurations, %d env-steps, no mismatch" % (n, steps))


if __name__ == "__main__":
    main()
