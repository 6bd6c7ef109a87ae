"""BinPackingVecEnv -- the reference's VecEnv surface over N bins living in B200 HBM.

Mirrors what `acktr.envs.make_vec_envs` returns (acktr/envs.py:77-118: ShmemVecEnv -> VecNormalize
pass-through -> VecPyTorch): `reset() -> obs float32[N,4LW]` on the device, `step(actions) ->
(obs, reward float32[N,1] on the CPU, done np.bool_[N], infos)`, `num_envs`, `observation_space`,
`action_space`, `step_async/step_wait`, `close`.  On top of that it keeps everything as device
tensors (`location_masks`, `reward_t`, `done_t`, `counter_t`, `ratio_t`) so that a trainer can drop the
host mask loop of main.py:163-169 and never leave the GPU.

All compute goes through the C ABI of libbpp_b200.so (include/bpp_b200.h); torch only provides
device memory and streams.
"""
import ctypes as C
import time

import numpy as np
import torch

from . import _cabi
from .spaces import Box, Discrete


def default_item_set(lo=(2, 2, 2), hi=(5, 5, 5)):
    """acktr/arguments.py:122-128: i-major product of --item-size-range."""
    return [(i, j, k) for i in range(lo[0], hi[0] + 1) for j in range(lo[1], hi[1] + 1)
            for k in range(lo[2], hi[2] + 1)]


class LazyInfos(object):
    """Sequence of per-env info dicts materialised on demand.

    The reference builds N dicts per step (envs/bpp0/bin3D.py:111,123-125 + Monitor's 'episode' entry,
    baselines/bench/monitor.py:58-75); main.py:159-162 only reads those of finished episodes.  At
    N = 2**19 a tuple of dicts per step is not an option, so dicts are built when indexed -- from the step's
    16-byte result records (include/bpp_b200.h bpp_result: placed volume and counter are integers, so `ratio`
    and the episode return are exact float64 values, like the reference's)."""

    def __init__(self, done, records, lwh, t0, act_len):
        self._done, self._rec = done, records               # np.bool_[N], int32[N,4] view of the pinned records
        self._lwh, self._t0, self._act_len = float(lwh), t0, act_len

    def __len__(self):
        return len(self._done)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        vol, counter = int(self._rec[i, 1]), int(self._rec[i, 2])
        info = {"counter": counter, "ratio": vol / self._lwh}              # space.py:146-151
        if self._done[i]:
            info["mask"] = np.ones(self._act_len)                           # bin3D.py:111
            # Monitor: r = round(sum(rewards), 6) with rewards 10*vol_i/LWH; l counts the failing step
            info["episode"] = {"r": round(10.0 * vol / self._lwh, 6), "l": counter + 1,
                               "t": round(time.time() - self._t0, 6)}
        return info

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def done_indices(self):
        return np.flatnonzero(self._done)


class _HostBuffer(object):
    """Pinned, device-mapped host memory on the NUMA node of the env's GPU (bpp_host_alloc), viewed as numpy / torch."""

    def __init__(self, lib, device_index, nbytes):
        self._lib, self.nbytes = lib, int(nbytes)
        ptr = C.c_void_p()
        _cabi.check(lib.bpp_host_alloc(int(device_index), self.nbytes, C.byref(ptr)))
        self.ptr = ptr.value
        self._raw = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        self.u8 = np.frombuffer(self._raw, dtype=np.uint8)

    def view(self, dtype, shape):
        return self.u8.view(dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.u8 = self._raw = None
            self._lib.bpp_host_free(C.c_void_p(self.ptr))
            self.ptr = None


class BinPackingVecEnv(object):
    """N independent bins on one GPU.

    Constructor keywords follow PackingGame (envs/bpp0/bin3D.py:10-12): container_size, box_set,
    enable_rotation, data_type in {'rs','cut1','cut2'}, data_name/test for dataset replay.
    """

    def __init__(self, num_envs, container_size=(10, 10, 10), box_set=None, enable_rotation=False,
                 data_type="rs", data_name=None, test=False, trajectories=None, first_traj=1,
                 mask_variant=2, seed=0, device=None, global_offset=0, global_n=None,
                 want_sample=False, host_infos=True):
        if not torch.cuda.is_available():
            raise RuntimeError("BinPackingVecEnv needs a CUDA device (libbpp_b200 has no CPU path)")
        self._lib = _cabi.lib()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.num_envs = int(num_envs)
        self.bin_size = tuple(int(v) for v in container_size)
        L, W, H = self.bin_size
        self.area = L * W
        self.can_rotate = bool(enable_rotation)
        self.act_len = self.area * (2 if self.can_rotate else 1)
        self.obs_len = self.area * 4
        self.observation_space = Box(0.0, H, (self.obs_len,))          # bin3D.py:41
        self.action_space = Discrete(self.act_len)                     # bin3D.py:40
        self.host_infos = host_infos
        self._keep = []
        cfg = _cabi.Config()
        cfg.abi_version = _cabi.ABI_VERSION
        cfg.device = self.device.index if self.device.index is not None else torch.cuda.current_device()
        cfg.n_bins = self.num_envs
        cfg.L, cfg.W, cfg.H = L, W, H
        cfg.rot = int(self.can_rotate)
        cfg.mask_variant = int(mask_variant)
        if test or trajectories is not None or data_type == "replay":
            if trajectories is None:
                trajectories = torch.load(data_name)                   # LoadBoxCreator, binCreator.py:47-49
            traj, traj_len = _pack_trajectories(trajectories)
            self._keep += [traj, traj_len]
            cfg.gen_kind = _cabi.GEN_REPLAY
            cfg.n_traj, cfg.traj_cap = traj.shape[0], traj.shape[1]
            cfg.traj_host = traj.ctypes.data
            cfg.traj_len_host = traj_len.ctypes.data
            cfg.first_traj = int(first_traj)
        else:
            cfg.gen_kind = _cabi.GEN_KINDS[data_type]
            if box_set is None:
                box_set = default_item_set()
            items = np.ascontiguousarray(np.asarray(box_set, dtype=np.uint8).reshape(-1, 3))
            self._keep.append(items)
            cfg.n_items = items.shape[0]
            cfg.item_set_host = items.ctypes.data
            if data_type == "cut2":                                    # given_bound = [box_set[0][0], box_set[-1][0]], bin3D.py:33
                for a in range(3):
                    cfg.lo[a], cfg.hi[a] = int(items[0][0]), int(items[-1][0])
            else:                                                      # low = box_set[0], up = box_set[-1], bin3D.py:25-30
                for a in range(3):
                    cfg.lo[a], cfg.hi[a] = int(items[0][a]), int(items[-1][a])
        cfg.seed = int(seed)
        cfg.global_offset = int(global_offset)
        cfg.global_n = int(global_n if global_n is not None else self.num_envs)
        self._cfg = cfg
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.bpp_create(C.byref(cfg), C.byref(h)))
        self._h = h
        N, dev = self.num_envs, self.device
        self.obs = torch.empty((N, self.obs_len), dtype=torch.float32, device=dev)
        self.location_masks = torch.empty((N, self.act_len), dtype=torch.float32, device=dev)
        self.reward_t = torch.zeros((N,), dtype=torch.float32, device=dev)
        self.done_t = torch.zeros((N,), dtype=torch.uint8, device=dev)
        self.counter_t = torch.zeros((N,), dtype=torch.int32, device=dev)
        self.ratio_t = torch.zeros((N,), dtype=torch.float32, device=dev)
        self.sample_t = torch.zeros((N,), dtype=torch.int64, device=dev) if want_sample else None
        # Host side of the reference-facing step(): pinned, device-mapped buffers next to this GPU.  Two result sets are
        # used alternately, so what step() returned stays valid until the call after the next one.
        di = cfg.device
        self._hbufs = [_HostBuffer(self._lib, di, 8 * N)] + [_HostBuffer(self._lib, di, 16 * N) for _ in range(2)] + \
                      [_HostBuffer(self._lib, di, N) for _ in range(2)]
        self._h_actions = torch.from_numpy(self._hbufs[0].view(np.int64, (N,)))
        self._h_rec = [self._hbufs[1 + k].view(np.int32, (N, 4)) for k in range(2)]
        self._h_reward = [torch.from_numpy(self._hbufs[1 + k].view(np.float32, (N, 4)))[:, 0:1] for k in range(2)]
        self._h_done = [self._hbufs[3 + k].view(np.bool_, (N,)) for k in range(2)]
        self._flip = 0
        self._dev_actions = None
        self._step_args = None
        self._pending = None
        self._t0 = time.time()
        self.venv = None        # get_vec_normalize(envs) walks .venv and must end at None (acktr/utils.py:96-102)

    # -- plumbing ------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _p(t):
        return C.c_void_p(0 if t is None else t.data_ptr())

    def close(self):
        if getattr(self, "_h", None):
            torch.cuda.synchronize(self.device)
            self._lib.bpp_destroy(self._h)
            self._h = None
            self._h_actions = self._h_rec = self._h_reward = self._h_done = None
            for hb in getattr(self, "_hbufs", ()):
                hb.free()
            self._hbufs = ()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- device-resident API ---------------------------------------------------------------------------
    def reset(self):
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.bpp_reset(self._h, self._p(self.obs), self._p(self.location_masks),
                                            self._p(self.sample_t), self._stream()))
        return self.obs

    def step_device(self, actions):
        """actions: int64 CUDA tensor [N] or [N,1].  Everything stays on the device; returns
        (obs, reward[N], done[N] uint8, location_masks)."""
        a = actions.reshape(-1)
        if a.dtype != torch.int64 or not a.is_cuda or not a.is_contiguous():
            a = a.to(device=self.device, dtype=torch.int64).contiguous()
        if a.numel() != self.num_envs:
            raise ValueError("expected %d actions, got %d" % (self.num_envs, a.numel()))
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.bpp_step(self._h, self._p(a), self._p(self.obs), self._p(self.location_masks),
                                           self._p(self.reward_t), self._p(self.done_t), self._p(self.counter_t),
                                           self._p(self.ratio_t), self._p(self.sample_t), self._stream()))
        return self.obs, self.reward_t, self.done_t, self.location_masks

    # -- reference-facing VecEnv API (acktr/envs.py:170-193) -------------------------------------------
    def step_async(self, actions):
        self._dev_actions = None
        if isinstance(actions, torch.Tensor):
            a = actions.reshape(-1)                                     # LongTensor[N,1] squeeze, envs.py:183-185
            if a.numel() != self.num_envs:
                raise ValueError("expected %d actions, got %d" % (self.num_envs, a.numel()))
            if a.is_cuda:
                # a policy that lives on this GPU: the kernel reads its actions where they are (the reference's
                # `.cpu()`, envs.py:186, has no counterpart here)
                self._dev_actions = a if (a.dtype == torch.int64 and a.is_contiguous()) else a.to(torch.int64).contiguous()
            else:
                self._h_actions.copy_(a)
        else:
            self._h_actions.copy_(torch.as_tensor(np.asarray(actions).reshape(-1), dtype=torch.int64))
        self._pending = True

    def step_wait(self):
        """-> (obs float32[N,4LW] on the device, reward float32[N,1] on the CPU, done np.bool_[N], infos).
        reward / done / infos are views of pinned host memory the kernel wrote over PCIe; they stay valid until the
        second step() after this one."""
        assert self._pending, "step_wait without step_async"
        self._pending = None
        k = self._flip
        self._flip ^= 1
        if self._step_args is None:            # the pointers of a step() call never change: build the ctypes values once
            self._step_args = [(C.c_void_p(self._hbufs[0].ptr), self._p(self.obs), self._p(self.location_masks),
                                C.c_void_p(self._hbufs[1 + j].ptr), C.c_void_p(self._hbufs[3 + j].ptr), self._p(self.sample_t))
                               for j in range(2)]
            self._lwh = float(np.prod(self.bin_size))
            self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        h_act, obs_p, mask_p, rec_p, done_p, sample_p = self._step_args[k]
        act = self._p(self._dev_actions) if self._dev_actions is not None else h_act
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        if torch.cuda.current_device() == self._dev_index:       # (the library selects the env's device itself)
            rc = self._lib.bpp_step_packed_host(self._h, act, obs_p, mask_p, rec_p, done_p, sample_p, stream)
        else:
            with torch.cuda.device(self.device):
                rc = self._lib.bpp_step_packed_host(self._h, act, obs_p, mask_p, rec_p, done_p, sample_p, stream)
        if rc != 0:
            _cabi.check(rc)
        self._dev_actions = None
        done = self._h_done[k]
        infos = LazyInfos(done, self._h_rec[k], self._lwh, self._t0, self.act_len) if self.host_infos else ()
        return self.obs, self._h_reward[k], done, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    # -- state / metrics -------------------------------------------------------------------------------
    def export_state(self):
        N, A, dev = self.num_envs, self.area, self.device
        plain = torch.empty((N, A), dtype=torch.int32, device=dev)
        box = torch.empty((N, 3), dtype=torch.int32, device=dev)
        counter = torch.empty((N,), dtype=torch.int32, device=dev)
        vol = torch.empty((N,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _cabi.check(self._lib.bpp_export_state(self._h, self._p(plain), self._p(box), self._p(counter),
                                                   self._p(vol), self._stream()))
        return plain, box, counter, vol

    def import_state(self, plain=None, box=None, counter=None, vol=None):
        def prep(t):
            return None if t is None else torch.as_tensor(t).to(device=self.device, dtype=torch.int32).contiguous()
        plain, box, counter, vol = prep(plain), prep(box), prep(counter), prep(vol)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.bpp_import_state(self._h, self._p(plain), self._p(box), self._p(counter),
                                                   self._p(vol), self._stream()))
            torch.cuda.current_stream(self.device).synchronize()

    def observe(self):
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.bpp_observe(self._h, self._p(self.obs), self._p(self.location_masks),
                                              self._stream()))
        return self.obs, self.location_masks

    def episode_items(self, bin_index=0, cap=4096):
        """cut1 / cut2: the item sequence of the episode `bin_index` is playing, as a list of (x, y, z)."""
        buf = np.zeros((cap, 3), dtype=np.uint8)
        n = C.c_int32(0)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.bpp_episode_items_host(self._h, int(bin_index), buf.ctypes.data_as(C.c_void_p), cap,
                                                        C.byref(n), self._stream()))
        return [tuple(int(v) for v in row) for row in buf[:min(n.value, cap)]]

    def state_dict(self):
        """Full env state (heightmaps, per-bin records incl. RNG counters, episode buffers, metrics) as one CUDA uint8
        tensor; `load_state_dict` resumes a rollout bit-exactly in an env created with the same arguments."""
        n = int(self._lib.bpp_state_bytes(self._h))
        blob = torch.empty((n,), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.bpp_save_state(self._h, self._p(blob), self._stream()))
        return {"blob": blob, "obs": self.obs.clone(), "location_masks": self.location_masks.clone(),
                "sample": None if self.sample_t is None else self.sample_t.clone()}

    def load_state_dict(self, sd):
        blob = sd["blob"].to(device=self.device, dtype=torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.bpp_load_state(self._h, self._p(blob), self._stream()))
        self.obs.copy_(sd["obs"])
        self.location_masks.copy_(sd["location_masks"])
        if self.sample_t is not None and sd.get("sample") is not None:
            self.sample_t.copy_(sd["sample"])

    def metrics(self):
        """Local episode accumulators: {'episodes','sum_counter','sum_vol','steps'} (python ints)."""
        out = (C.c_uint64 * 4)()
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.bpp_metrics_host(self._h, out, self._stream()))
        return {"episodes": int(out[0]), "sum_counter": int(out[1]), "sum_vol": int(out[2]), "steps": int(out[3])}


def _pack_trajectories(trajectories):
    """list[list[[x,y,z]]] (dataset/*.pt) or (uint8[T,K,3], int32[T]) -> padded arrays."""
    if isinstance(trajectories, tuple) and len(trajectories) == 2:
        traj = np.ascontiguousarray(trajectories[0], dtype=np.uint8)
        traj_len = np.ascontiguousarray(trajectories[1], dtype=np.int32)
        return traj, traj_len
    K = max(len(t) for t in trajectories)
    traj = np.zeros((len(trajectories), K, 3), dtype=np.uint8)
    traj_len = np.zeros(len(trajectories), dtype=np.int32)
    for i, t in enumerate(trajectories):
        if len(t):
            traj[i, :len(t)] = np.asarray(t, dtype=np.uint8).reshape(-1, 3)
        traj_len[i] = len(t)
    return traj, traj_len
