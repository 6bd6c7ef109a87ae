#!/usr/bin/env python
"""A/B timing of kernel build variants in ONE process (one torch import, same box, same clocks).

    python online-3d-bpp-drl_b200/build.py --out=.../variants/libbpp_b200_full.so -DBPP_PACKED_SWEEP=1 -DBPP_LANE_CHECK=1
    python tools/ab_bench.py --libs default=online-3d-bpp-drl_b200/b200bpp/libbpp_b200.so full=.../libbpp_b200_full.so \
        --workloads target target_rot cfg3 --steps 300 --rounds 3

Every library is loaded with its own ctypes handle and driven through the C ABI exactly like bench.py's device-resident
leg (bpp_step, actions = the in-kernel sampled valid action, CUDA events on the launching stream).  Rounds interleave the
libraries so that clock drift hits all of them alike; the minimum over rounds is reported.  The episode metrics after
the same number of steps from the same seed must be identical across libraries: a full-size parity cross-check for free.
Not a bench value: bench.py is the contract; this only ranks variants.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "online-3d-bpp-drl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np   # noqa: E402


def bind(path):
    from b200bpp import _cabi
    l = C.CDLL(os.path.abspath(path))
    for name, (res, args) in _cabi.SYMBOLS.items():
        f = getattr(l, name)
        f.restype, f.argtypes = res, args
    assert l.bpp_abi_version() == _cabi.ABI_VERSION
    return l


class Env(object):
    """The few lines of BinPackingVecEnv needed here, over an explicitly chosen library."""

    def __init__(self, lib, bins, size, rot, gen, seed, dev):
        import torch
        from b200bpp import _cabi
        from b200bpp.vec_env import default_item_set
        self.lib, self.N, self.dev = lib, bins, dev
        L, W, H = size
        A = L * W
        cfg = _cabi.Config()
        cfg.abi_version, cfg.device, cfg.n_bins = _cabi.ABI_VERSION, dev.index, bins
        cfg.L, cfg.W, cfg.H, cfg.rot, cfg.mask_variant = L, W, H, int(rot), 2
        self._keep = []
        if gen == "cut2_replay":
            d = np.load(os.path.join(ROOT, "tests", "golden", "datasets.npz"))
            traj = np.ascontiguousarray(d["cut_2"], dtype=np.uint8)
            tl = np.ascontiguousarray(d["cut_2_len"], dtype=np.int32)
            self._keep += [traj, tl]
            cfg.gen_kind, cfg.n_traj, cfg.traj_cap = _cabi.GEN_REPLAY, traj.shape[0], traj.shape[1]
            cfg.traj_host, cfg.traj_len_host, cfg.first_traj = traj.ctypes.data, tl.ctypes.data, 1
        else:
            items = np.ascontiguousarray(np.asarray(default_item_set(), dtype=np.uint8))
            self._keep.append(items)
            cfg.gen_kind, cfg.n_items, cfg.item_set_host = _cabi.GEN_KINDS[gen], items.shape[0], items.ctypes.data
            for a in range(3):
                cfg.lo[a], cfg.hi[a] = 2, 5
        cfg.seed, cfg.global_offset, cfg.global_n = seed, 0, bins
        h = C.c_void_p()
        rc = lib.bpp_create(C.byref(cfg), C.byref(h))
        assert rc == 0, lib.bpp_last_error()
        self.h = h
        f32, dv = torch.float32, dev
        self.obs = torch.empty((bins, 4 * A), dtype=f32, device=dv)
        self.mask = torch.empty((bins, A * (2 if rot else 1)), dtype=f32, device=dv)
        self.reward = torch.empty(bins, dtype=f32, device=dv)
        self.done = torch.empty(bins, dtype=torch.uint8, device=dv)
        self.counter = torch.empty(bins, dtype=torch.int32, device=dv)
        self.ratio = torch.empty(bins, dtype=f32, device=dv)
        self.act = [torch.zeros(bins, dtype=torch.int64, device=dv) for _ in range(2)]
        self.k = 0
        self.stream = torch.cuda.current_stream(dev)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr())

    def reset(self):
        rc = self.lib.bpp_reset(self.h, self._p(self.obs), self._p(self.mask), self._p(self.act[0]),
                                C.c_void_p(self.stream.cuda_stream))
        assert rc == 0, self.lib.bpp_last_error()
        self.k = 0

    def step(self):
        src, dst = self.act[self.k & 1], self.act[(self.k + 1) & 1]
        rc = self.lib.bpp_step(self.h, self._p(src), self._p(self.obs), self._p(self.mask), self._p(self.reward),
                               self._p(self.done), self._p(self.counter), self._p(self.ratio), self._p(dst),
                               C.c_void_p(self.stream.cuda_stream))
        assert rc == 0, self.lib.bpp_last_error()
        self.k += 1

    def metrics(self):
        out = (C.c_uint64 * 4)()
        assert self.lib.bpp_metrics_host(self.h, out, C.c_void_p(self.stream.cuda_stream)) == 0
        return [int(v) for v in out]

    def close(self):
        import torch
        torch.cuda.synchronize(self.dev)
        self.lib.bpp_destroy(self.h)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--libs", nargs="+", required=True, help="name=path ...")
    ap.add_argument("--workloads", nargs="+", default=["target"])
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--bw", action="store_true", help="also time write-only / read-only / copy streams (1 GiB, torch kernels)")
    args = ap.parse_args()
    import torch
    import bench
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    libs = [(s.split("=", 1)[0], bind(s.split("=", 1)[1])) for s in args.libs]
    results = {}
    if args.bw:
        # what a write-dominated kernel can hope for: the step kernel writes 13x more than it reads, while the roofline
        # denominator (MEASURED_PEAKS.json) is a copy, i.e. half reads
        n = 1 << 28
        a = torch.empty(n, dtype=torch.float32, device=dev)
        b = torch.empty(n, dtype=torch.float32, device=dev)
        a.fill_(1.0); b.fill_(2.0)
        st = torch.cuda.current_stream(dev)

        def best_ms(fn, reps=6):
            t = float("inf")
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st); fn(); e1.record(st)
                torch.cuda.synchronize(dev)
                t = min(t, e0.elapsed_time(e1))
            return t
        gb = n * 4 / 1e9
        bw = {"write_only_GBps": gb / (best_ms(lambda: a.fill_(3.0)) * 1e-3),
              "read_only_GBps": gb / (best_ms(lambda: a.sum()) * 1e-3),
              "copy_GBps": 2 * gb / (best_ms(lambda: b.copy_(a)) * 1e-3), "bytes": n * 4}
        print(json.dumps({"bandwidth": bw}), flush=True)
        del a, b
    for wl in args.workloads:
        if ":" in wl:                                   # ad hoc: BINS:ROT, 10x10x10 bins, rs generator
            bins, size, rot, gen = int(wl.split(":")[0]), (10, 10, 10), wl.split(":")[1] == "1", "rs"
        else:
            bins, size, rot, gen, desc = bench.WORKLOADS[wl]
        envs = []
        for name, lib in libs:
            e = Env(lib, bins, size, rot, gen, args.seed, dev)
            e.reset()
            for _ in range(args.warmup):
                e.step()
            envs.append((name, e))
        torch.cuda.synchronize(dev)
        best = {name: float("inf") for name, _ in envs}
        for _ in range(args.rounds):
            for name, e in envs:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(e.stream)
                for _ in range(args.steps):
                    e.step()
                e1.record(e.stream)
                torch.cuda.synchronize(dev)
                best[name] = min(best[name], e0.elapsed_time(e1) / args.steps)
        met = {name: e.metrics() for name, e in envs}
        for _, e in envs:
            e.close()
        ref = met[envs[0][0]]
        results[wl] = {name: {"ms_per_step": best[name], "env_steps_per_s": bins / (best[name] * 1e-3),
                              "roofline_frac": bins * bench.algorithmic_bytes(size, rot) / (best[name] * 1e-3) / 1e9
                              / bench.hbm_peak()[0],
                              "metrics": met[name], "metrics_equal_to_first": met[name] == ref,
                              "build": dict(libs)[name].bpp_build_info().decode()} for name, _ in envs}
        print(json.dumps({wl: results[wl]}), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
