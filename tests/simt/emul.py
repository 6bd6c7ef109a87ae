"""Python front end of the CPU interpretation of the CUDA kernel sources (tests/simt/simt_emul.h).  TESTS ONLY.

`EmulEnv` plays the role of bpp_api.cu + BinPackingVecEnv for an env whose "device" memory is numpy arrays: it builds
the same bpp_config, derives the kernel parameters with the product's own `derive_params`, and calls the product's
`launch_step` / `launch_gen` / `launch_mask_*` compiled by g++ against the SIMT shim.  Interface = oracle.OracleEnv,
so the parity tests read alike.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from b200bpp import _cabi

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.abspath(os.path.join(_HERE, "..", ".."))
_CSRC = os.path.join(_ROOT, "online-3d-bpp-drl_b200", "csrc")
_DEPS = [os.path.join(_HERE, f) for f in ("simt_emul.h", "emul_kernels.cpp")] + \
        [os.path.join(_CSRC, f) for f in ("bpp_common.cuh", "bpp_launch.h", "bpp_params.h", "bpp_step.cu", "bpp_gen.cu",
                                          "bpp_mask.cu")] + [os.path.join(_ROOT, "include", "bpp_b200.h")]


def _source_digest(defines):
    """Content hash of everything the build depends on (mtimes lie after a `mv` or a `git checkout`)."""
    import hashlib
    h = hashlib.sha1(("|".join(defines)).encode())
    for d in _DEPS:
        h.update(open(d, "rb").read())
    return h.hexdigest()[:12]


def build(defines=(), out=None):
    """g++ -> tests/simt/_build/libbpp_emul_<variant>_<digest>.so (git-ignored).  `defines` selects kernel build switches
    like build.py -D...; the digest of the sources is part of the file name, so a stale build can never be picked up."""
    tag = "_".join(d.replace("=", "").replace("BPP_", "").lower() for d in defines) or "default"
    out = out or os.path.join(_HERE, "_build", "libbpp_emul_%s_%s.so" % (tag[:80], _source_digest(defines)))
    if os.path.exists(out):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    for old in os.listdir(os.path.dirname(out)):              # earlier builds of this variant
        if old.startswith("libbpp_emul_%s_" % tag[:80]) and old.endswith(".so"):
            os.remove(os.path.join(os.path.dirname(out), old))
    # -fsanitize=alignment: a 128-bit access the GPU would fault on ("misaligned address") aborts here too
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-fsanitize=alignment", "-fno-sanitize-recover=alignment", "-I", _HERE] + ["-D" + d for d in defines] + \
          ["-o", out + ".tmp", os.path.join(_HERE, "emul_kernels.cpp")]
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    os.replace(out + ".tmp", out)
    return out


class _Buffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("hm", "sc", "item_set", "traj", "traj_len", "metrics", "ep_items", "ep_len2",
                                          "gen_list", "gen_count", "gen_scratch")]


class _IO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("actions", "obs", "mask", "reward", "done", "counter", "ratio", "sample", "packed")] + \
               [("n_steps", C.c_int32), ("actions32", C.c_void_p), ("actions32_stride", C.c_int32)]


_libs = {}


def lib(defines=()):
    key = tuple(defines)
    if key not in _libs:
        l = C.CDLL(build(key))
        l.emul_env_call.restype = C.c_int
        l.emul_env_call.argtypes = [C.POINTER(_cabi.Config), C.POINTER(_Buffers), C.POINTER(_IO)] + [C.c_int] * 4
        l.emul_sizes.restype = None
        l.emul_sizes.argtypes = [C.POINTER(_cabi.Config), C.c_void_p]
        l.emul_mask_from_obs.restype = C.c_int
        l.emul_mask_from_obs.argtypes = [C.c_void_p] + [C.c_int] * 6 + [C.c_void_p, C.c_int]
        l.emul_mask_from_plain.restype = C.c_int
        l.emul_mask_from_plain.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p, C.c_int]
        l.emul_check_box.restype = C.c_int
        l.emul_check_box.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]
        l.emul_export_state.restype = C.c_int
        l.emul_export_state.argtypes = [C.POINTER(_cabi.Config), C.POINTER(_Buffers)] + [C.c_void_p] * 4
        l.emul_import_state.restype = C.c_int
        l.emul_import_state.argtypes = [C.POINTER(_cabi.Config), C.POINTER(_Buffers)] + [C.c_void_p] * 4
        l.emul_collectives.restype = C.c_uint64
        l.emul_debug_draws.restype = C.c_int
        l.emul_debug_draws.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, C.c_void_p]
        l.emul_set_lane_order.restype = None
        l.emul_set_lane_order.argtypes = [C.c_int]
        _libs[key] = l
    return _libs[key]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _aligned(shape, dtype, fill=0):
    """16-byte aligned array (the kernels use 128-bit accesses on rows)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    raw = np.zeros(n + 128, dtype=np.uint8)       # >= 64 readable bytes after the array (kHmSlack)
    off = (-raw.ctypes.data) % 64
    a = raw[off:off + n].view(dtype).reshape(shape)
    a[...] = fill
    return a


def pack_items(items):
    it = np.asarray(items, dtype=np.uint32).reshape(-1, 3)
    return np.ascontiguousarray(it[:, 0] | (it[:, 1] << 8) | (it[:, 2] << 16), dtype=np.uint32)


class EmulEnv(object):
    def __init__(self, n_bins, container_size=(10, 10, 10), rot=False, mask_variant=2, gen="rs", item_set=None,
                 traj=None, traj_len=None, first_traj=1, lo=(2, 2, 2), hi=(5, 5, 5), seed=0, global_offset=0,
                 global_n=None, num_sms=2, blocks_per_sm=1, defines=()):
        self._lib = lib(defines)
        self.N = int(n_bins)
        self.L, self.W, self.H = (int(v) for v in container_size)
        self.A, self.R = self.L * self.W, (2 if rot else 1)
        self.num_sms, self.blocks_per_sm = int(num_sms), int(blocks_per_sm)
        cfg = _cabi.Config()
        cfg.abi_version, cfg.device, cfg.n_bins = _cabi.ABI_VERSION, 0, self.N
        cfg.L, cfg.W, cfg.H, cfg.rot, cfg.mask_variant = self.L, self.W, self.H, int(bool(rot)), int(mask_variant)
        cfg.gen_kind = _cabi.GEN_KINDS["replay" if gen == "replay" else gen]
        self._keep = []
        b = _Buffers()
        if gen == "rs":
            if item_set is None:
                item_set = [(i, j, k) for i in range(lo[0], hi[0] + 1) for j in range(lo[1], hi[1] + 1)
                            for k in range(lo[2], hi[2] + 1)]
            items = pack_items(item_set)
            cfg.n_items = items.shape[0]
            self._keep.append(items)
            b.item_set = items.ctypes.data
        elif gen == "replay":
            traj = np.asarray(traj, dtype=np.uint32)
            traj_len = np.ascontiguousarray(traj_len, dtype=np.int32)
            packed = traj[:, :, 0] | (traj[:, :, 1] << 8) | (traj[:, :, 2] << 16)
            packed = np.where(np.arange(traj.shape[1])[None, :] < traj_len[:, None], packed, 0)
            packed = np.ascontiguousarray(packed, dtype=np.uint32)
            cfg.n_traj, cfg.traj_cap, cfg.first_traj = traj.shape[0], traj.shape[1], int(first_traj)
            self._keep += [packed, traj_len]
            b.traj, b.traj_len = packed.ctypes.data, traj_len.ctypes.data
        for a in range(3):
            cfg.lo[a], cfg.hi[a] = int(lo[a]), int(hi[a])
        if gen == "cut2":                       # vec_env.py / bin3D.py:33: one range for all axes
            for a in range(3):
                cfg.lo[a], cfg.hi[a] = int(lo[0]), int(hi[0])
        cfg.seed, cfg.global_offset = int(seed), int(global_offset)
        cfg.global_n = int(global_n if global_n is not None else self.N)
        self.cfg = cfg
        sz = np.zeros(6, dtype=np.int64)
        self._lib.emul_sizes(C.byref(cfg), _ptr(sz))
        self.A_pad, self.M_pad, self.ep_cap, self.scratch_stride, sc_bytes, self.smem = (int(v) for v in sz)
        N = self.N
        self.hm = _aligned((N, self.A_pad), np.uint8)
        self.sc = _aligned((N, sc_bytes), np.uint8)
        self.metrics = np.zeros(4, dtype=np.uint64)
        b.hm, b.sc, b.metrics = self.hm.ctypes.data, self.sc.ctypes.data, self.metrics.ctypes.data
        if gen in ("cut1", "cut2"):
            self.ep_items = _aligned((2, N, self.ep_cap), np.uint32)
            self.ep_len2 = np.zeros((2, N), dtype=np.int32)
            self.gen_list = np.zeros(N, dtype=np.int32)
            self.gen_count = np.zeros(1, dtype=np.int32)
            self.gen_scratch = _aligned((N, self.scratch_stride), np.uint8, fill=0xCD)
            b.ep_items, b.ep_len2 = self.ep_items.ctypes.data, self.ep_len2.ctypes.data
            b.gen_list, b.gen_count = self.gen_list.ctypes.data, self.gen_count.ctypes.data
            b.gen_scratch = self.gen_scratch.ctypes.data
        self.buf = b
        self.obs = _aligned((N, 4 * self.A), np.float32, fill=-1)
        self.mask = _aligned((N, self.A * self.R), np.float32, fill=-1)
        self.reward = np.zeros(N, dtype=np.float32)
        self.done = np.zeros(N, dtype=np.uint8)
        self.counter = np.zeros(N, dtype=np.int32)
        self.ratio = np.zeros(N, dtype=np.float32)
        self.sample = np.zeros(N, dtype=np.int64)
        self.was_reset = False
        self.steps = 0

    def _call(self, mode, actions=None, want_obs=True, want_mask=True, want_sample=True, scalars=True, packed=False,
              n_steps=1, actions32=None, actions32_stride=1):
        """packed: results as 16-byte records (self.packed: reward f32 | vol i32 | counter i32 | next action i32) instead of
        reward / counter / ratio / sample; n_steps: env steps in one launch (later steps take the sampled action)."""
        io = _IO()
        io.n_steps = int(n_steps)
        if actions32 is not None:       # int32 actions with a stride (e.g. the next_action column of the previous records)
            io.actions32, io.actions32_stride = actions32.ctypes.data, int(actions32_stride)
        if packed and mode == 0:
            if getattr(self, "packed", None) is None:
                self.packed = _aligned((self.N, 4), np.int32)
            io.packed = self.packed.ctypes.data
            want_sample = want_sample and packed != "only"
        io.actions = None if actions is None else actions.ctypes.data
        io.obs = self.obs.ctypes.data if want_obs else None
        io.mask = self.mask.ctypes.data if want_mask else None
        io.sample = self.sample.ctypes.data if want_sample else None
        if mode == 0 and scalars:
            io.done = self.done.ctypes.data
            if not (packed and mode == 0):
                io.reward = self.reward.ctypes.data
                io.counter, io.ratio = self.counter.ctypes.data, self.ratio.ctypes.data
        rc = self._lib.emul_env_call(C.byref(self.cfg), C.byref(self.buf), C.byref(io), mode, int(self.was_reset),
                                     self.num_sms, self.blocks_per_sm)
        assert rc == 0

    def reset(self, **kw):
        self._call(1, **kw)
        self.was_reset = True
        return self.obs, self.mask

    def step(self, actions, **kw):
        a = np.ascontiguousarray(actions, dtype=np.int64).reshape(-1)
        assert a.shape[0] == self.N and self.was_reset
        self._call(0, a, **kw)
        self.steps += self.N
        return self.obs, self.reward, self.done, self.counter, self.ratio

    def observe(self):
        self._call(2, want_sample=False)
        return self.obs, self.mask

    def get_state(self):
        plain = np.empty((self.N, self.A), dtype=np.int32)
        box = np.empty((self.N, 3), dtype=np.int32)
        counter = np.empty(self.N, dtype=np.int32)
        vol = np.empty(self.N, dtype=np.int32)
        assert self._lib.emul_export_state(C.byref(self.cfg), C.byref(self.buf), _ptr(plain), _ptr(box), _ptr(counter),
                                           _ptr(vol)) == 0
        return plain, box, counter, vol

    def set_state(self, plain=None, box=None, counter=None, vol=None):
        def prep(a):
            return None if a is None else np.ascontiguousarray(a, dtype=np.int32)
        plain, box, counter, vol = prep(plain), prep(box), prep(counter), prep(vol)
        assert self._lib.emul_import_state(C.byref(self.cfg), C.byref(self.buf), _ptr(plain), _ptr(box), _ptr(counter),
                                           _ptr(vol)) == 0
        self.was_reset = True

    def metrics_dict(self):
        return {"episodes": int(self.metrics[0]), "sum_counter": int(self.metrics[1]), "sum_vol": int(self.metrics[2]),
                "steps": int(self.steps)}


def mask_from_obs(obs, size, rot=False, variant=2, grid=2, defines=()):
    obs = np.ascontiguousarray(obs, dtype=np.float32)
    L, W, H = size
    out = np.full((obs.shape[0], L * W * (2 if rot else 1)), -1, dtype=np.float32)
    assert lib(defines).emul_mask_from_obs(_ptr(obs), obs.shape[0], L, W, H, int(rot), variant, _ptr(out), grid) == 0
    return out


def mask_from_plain(plain, box, size, rot=False, variant=2, grid=2, defines=()):
    plain = np.ascontiguousarray(plain, dtype=np.int32)
    box = np.ascontiguousarray(box, dtype=np.int32)
    L, W, H = size
    out = np.full((plain.shape[0], L * W * (2 if rot else 1)), -1, dtype=np.float32)
    assert lib(defines).emul_mask_from_plain(_ptr(plain), _ptr(box), plain.shape[0], L, W, H, int(rot), variant, _ptr(out),
                                             grid) == 0
    return out


def check_box(plain, query, size, variant=1, defines=()):
    plain = np.ascontiguousarray(plain, dtype=np.int32)
    query = np.ascontiguousarray(query, dtype=np.int32)
    L, W, H = size
    out = np.full(plain.shape[0], -7, dtype=np.int32)
    assert lib(defines).emul_check_box(_ptr(plain), _ptr(query), plain.shape[0], L, W, H, variant, _ptr(out)) == 0
    return out


def set_lane_order(order, defines=()):
    """0: lanes of a warp resume in order 0..31 between collectives, 1: 31..0, 2: odd lanes first."""
    lib(defines).emul_set_lane_order(int(order))


def debug_draws(seed, domain, aux, gid, n, defines=()):
    """bpp_debug_draws_kernel: the first n draws of Philox stream (seed, domain, aux, gid) as the kernels compute them."""
    out = np.zeros(n, dtype=np.uint32)
    assert lib(defines).emul_debug_draws(int(seed), int(domain), int(aux), int(gid), int(n), _ptr(out)) == 0
    return out
