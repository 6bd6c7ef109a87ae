"""ctypes/numpy front-end of the CPU oracle (oracle/bpp_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbpp_oracle.so")

GEN_RS, GEN_REPLAY, GEN_CUT1, GEN_CUT2 = 0, 1, 2, 3
GEN_KINDS = {"rs": GEN_RS, "replay": GEN_REPLAY, "cut1": GEN_CUT1, "cut2": GEN_CUT2}
MASK_V1, MASK_V2 = 1, 2


def build(force: bool = False) -> str:
    """Compile the oracle with the system gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "bpp_oracle.c")
    hdr = os.path.join(_HERE, "bpp_oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    subprocess.check_call(["make", "-s", "-B", "-C", _HERE, "libbpp_oracle.so"])
    return _LIB_PATH


class _Config(C.Structure):
    _fields_ = [
        ("n_bins", C.c_int32), ("L", C.c_int32), ("W", C.c_int32), ("H", C.c_int32),
        ("rot", C.c_int32), ("mask_variant", C.c_int32), ("gen_kind", C.c_int32),
        ("n_items", C.c_int32), ("item_set", C.c_void_p),
        ("T", C.c_int32), ("K", C.c_int32), ("traj", C.c_void_p), ("traj_len", C.c_void_p),
        ("first_traj", C.c_int32), ("lo", C.c_int32 * 3), ("hi", C.c_int32 * 3),
        ("seed", C.c_uint64), ("global_offset", C.c_int64), ("global_n", C.c_int64),
        ("n_threads", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        i32, vp = C.c_int32, C.c_void_p
        _lib.orc_check_box_v1.restype = i32
        _lib.orc_check_box_v1.argtypes = [vp] + [i32] * 8
        _lib.orc_check_box_v2.restype = i32
        _lib.orc_check_box_v2.argtypes = [vp] + [i32] * 8
        _lib.orc_mask.restype = None
        _lib.orc_mask.argtypes = [vp] + [i32] * 8 + [vp]
        _lib.orc_philox4x32_10.restype = None
        _lib.orc_philox4x32_10.argtypes = [vp, vp, vp]
        _lib.orc_draw.restype = C.c_uint32
        _lib.orc_draw.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32]
        for f in (_lib.orc_gen_cut1, _lib.orc_gen_cut2):
            f.restype = i32
            f.argtypes = [C.POINTER(_Config), C.c_uint64, C.c_uint32, vp, vp, i32]
        _lib.orc_create.restype = vp
        _lib.orc_create.argtypes = [C.POINTER(_Config)]
        _lib.orc_destroy.restype = None
        _lib.orc_destroy.argtypes = [vp]
        _lib.orc_reset.restype = None
        _lib.orc_reset.argtypes = [vp] * 4
        _lib.orc_step.restype = None
        _lib.orc_step.argtypes = [vp] * 9
        _lib.orc_observe.restype = None
        _lib.orc_observe.argtypes = [vp] * 3
        _lib.orc_get_state.restype = None
        _lib.orc_get_state.argtypes = [vp] * 5
        _lib.orc_set_state.restype = None
        _lib.orc_set_state.argtypes = [vp] * 5
        _lib.orc_metrics.restype = None
        _lib.orc_metrics.argtypes = [vp, vp]
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_item_set(lo=(2, 2, 2), hi=(5, 5, 5)) -> np.ndarray:
    """acktr/arguments.py:122-128: i-major product of the per-axis ranges."""
    return np.array([(i, j, k) for i in range(lo[0], hi[0] + 1) for j in range(lo[1], hi[1] + 1)
                     for k in range(lo[2], hi[2] + 1)], dtype=np.uint8)


def check_box(plain, x, y, lx, ly, z, container_size, variant=MASK_V2) -> int:
    plain = np.ascontiguousarray(plain, dtype=np.int32)
    L, W, H = (int(v) for v in container_size)
    f = lib().orc_check_box_v1 if variant == MASK_V1 else lib().orc_check_box_v2
    return int(f(_ptr(plain), L, W, H, int(x), int(y), int(lx), int(ly), int(z)))


def mask(plain, box, container_size, rot=False, variant=MASK_V2) -> np.ndarray:
    plain = np.ascontiguousarray(plain, dtype=np.int32)
    L, W, H = (int(v) for v in container_size)
    out = np.empty(L * W * (2 if rot else 1), dtype=np.int32)
    lib().orc_mask(_ptr(plain), L, W, H, int(box[0]), int(box[1]), int(box[2]), int(bool(rot)),
                   int(variant), _ptr(out))
    return out


def philox(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.empty(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_ptr(c), _ptr(k), _ptr(o))
    return o


def draw(seed, domain, aux, gid, n) -> int:
    return int(lib().orc_draw(seed, domain, aux, gid, n))


def _make_config(n_bins, container_size, rot, mask_variant, gen, item_set, traj, traj_len,
                 first_traj, lo, hi, seed, global_offset, global_n, n_threads, keep):
    cfg = _Config()
    cfg.n_bins = int(n_bins)
    cfg.L, cfg.W, cfg.H = (int(v) for v in container_size)
    cfg.rot = int(bool(rot))
    cfg.mask_variant = int(mask_variant)
    cfg.gen_kind = GEN_KINDS[gen] if isinstance(gen, str) else int(gen)
    if item_set is not None:
        item_set = np.ascontiguousarray(item_set, dtype=np.uint8).reshape(-1, 3)
        keep.append(item_set)
        cfg.n_items = item_set.shape[0]
        cfg.item_set = item_set.ctypes.data
    if traj is not None:
        traj = np.ascontiguousarray(traj, dtype=np.uint8)
        traj_len = np.ascontiguousarray(traj_len, dtype=np.int32)
        keep.extend([traj, traj_len])
        cfg.T, cfg.K = traj.shape[0], traj.shape[1]
        cfg.traj = traj.ctypes.data
        cfg.traj_len = traj_len.ctypes.data
    cfg.first_traj = int(first_traj)
    for a in range(3):
        cfg.lo[a] = int(lo[a])
        cfg.hi[a] = int(hi[a])
    cfg.seed = int(seed)
    cfg.global_offset = int(global_offset)
    cfg.global_n = int(global_n if global_n is not None else n_bins)
    cfg.n_threads = int(n_threads)
    return cfg


def gen_episode(kind, container_size, lo, hi, seed, gid, episode, rot=False):
    """One cut1/cut2 episode -> (items[k,3], positions[k,3])."""
    keep = []
    cfg = _make_config(1, container_size, rot, MASK_V2, kind, None, None, None, 1, lo, hi, seed, 0, 1, 1, keep)
    cap = (cfg.L // lo[0]) * (cfg.W // lo[1]) * (cfg.H // lo[2]) + 8
    items = np.zeros((cap, 3), dtype=np.uint8)
    pos = np.zeros((cap, 3), dtype=np.uint8)
    f = lib().orc_gen_cut1 if kind == "cut1" else lib().orc_gen_cut2
    n = f(C.byref(cfg), int(gid), int(episode), _ptr(items), _ptr(pos), cap)
    return items[:n].copy(), pos[:n].copy()


class OracleEnv:
    """N independent bins stepped on the CPU with the reference's semantics."""

    def __init__(self, n_bins, container_size=(10, 10, 10), rot=False, mask_variant=MASK_V2,
                 gen="rs", item_set=None, traj=None, traj_len=None, first_traj=1,
                 lo=(2, 2, 2), hi=(5, 5, 5), seed=0, global_offset=0, global_n=None, n_threads=1):
        self._keep = []
        if gen == "rs" and item_set is None:
            item_set = default_item_set(lo, hi)
        self.cfg = _make_config(n_bins, container_size, rot, mask_variant, gen, item_set, traj,
                                traj_len, first_traj, lo, hi, seed, global_offset, global_n,
                                n_threads, self._keep)
        self.N = int(n_bins)
        self.L, self.W, self.H = self.cfg.L, self.cfg.W, self.cfg.H
        self.A = self.L * self.W
        self.R = 2 if rot else 1
        self._h = lib().orc_create(C.byref(self.cfg))
        N, A, R = self.N, self.A, self.R
        self.obs = np.zeros((N, 4 * A), dtype=np.float32)
        self.mask = np.zeros((N, A * R), dtype=np.float32)
        self.reward = np.zeros(N, dtype=np.float32)
        self.done = np.zeros(N, dtype=np.uint8)
        self.counter = np.zeros(N, dtype=np.int32)
        self.ratio = np.zeros(N, dtype=np.float32)
        self.sample = np.zeros(N, dtype=np.int64)

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, want_obs=True):
        lib().orc_reset(self._h, _ptr(self.obs) if want_obs else None, _ptr(self.mask), _ptr(self.sample))
        return self.obs, self.mask

    def step(self, actions, want_obs=True):
        a = np.ascontiguousarray(actions, dtype=np.int64).reshape(-1)
        assert a.shape[0] == self.N
        lib().orc_step(self._h, _ptr(a), _ptr(self.obs) if want_obs else None, _ptr(self.mask),
                       _ptr(self.reward), _ptr(self.done), _ptr(self.counter), _ptr(self.ratio),
                       _ptr(self.sample))
        return self.obs, self.reward, self.done, self.counter, self.ratio

    def observe(self):
        lib().orc_observe(self._h, _ptr(self.obs), _ptr(self.mask))
        return self.obs, self.mask

    def get_state(self):
        plain = np.empty((self.N, self.A), dtype=np.int32)
        box = np.empty((self.N, 3), dtype=np.int32)
        counter = np.empty(self.N, dtype=np.int32)
        vol = np.empty(self.N, dtype=np.int32)
        lib().orc_get_state(self._h, _ptr(plain), _ptr(box), _ptr(counter), _ptr(vol))
        return plain, box, counter, vol

    def set_state(self, plain=None, box=None, counter=None, vol=None):
        def prep(a):
            return None if a is None else np.ascontiguousarray(a, dtype=np.int32)
        plain, box, counter, vol = prep(plain), prep(box), prep(counter), prep(vol)
        lib().orc_set_state(self._h, _ptr(plain), _ptr(box), _ptr(counter), _ptr(vol))

    def metrics(self):
        out = np.zeros(4, dtype=np.uint64)
        lib().orc_metrics(self._h, _ptr(out))
        return {"episodes": int(out[0]), "sum_counter": int(out[1]), "sum_vol": int(out[2]),
                "steps": int(out[3])}
