"""Drop-ins for the training-side mask helpers of acktr/utils.py:8-94, computed on the GPU.

Same names, argument meaning and return types as the reference; `*_batch` variants take a whole
[N, 4LW] observation tensor and return a device tensor (what main.py:163-169 should call instead of
its per-row host loop).
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _device_of(t):
    return t.device if isinstance(t, torch.Tensor) and t.is_cuda else torch.device("cuda", torch.cuda.current_device())


def mask_batch(observations, container_size, enable_rotation=False, variant=_cabi.MASK_V2):
    """observations: float32 [n, 4LW] (torch, any device, or numpy) -> float32 CUDA tensor [n, LW(1+rot)]."""
    L, W, H = (int(v) for v in container_size)
    dev = _device_of(observations)
    obs = torch.as_tensor(observations).to(device=dev, dtype=torch.float32).reshape(-1, 4 * L * W).contiguous()
    n = obs.shape[0]
    out = torch.empty((n, L * W * (2 if enable_rotation else 1)), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().bpp_mask_from_obs(C.c_void_p(obs.data_ptr()), n, L, W, H, int(bool(enable_rotation)),
                                                  int(variant), C.c_void_p(out.data_ptr()), _stream(dev)))
    return out


def get_possible_position(observation, container_size):
    """acktr/utils.py:37-62: list[int] of length LW."""
    m = mask_batch(torch.as_tensor(observation).reshape(1, -1), container_size, False)
    return m[0].to(torch.int32).cpu().numpy().tolist()


def get_rotation_mask(observation, container_size):
    """acktr/utils.py:64-94: np.int32[2LW]."""
    m = mask_batch(torch.as_tensor(observation).reshape(1, -1), container_size, True)
    return m[0].to(torch.int32).cpu().numpy()


def check_box(plain, x, y, lx, ly, z, container_size, variant=_cabi.MASK_V2):
    """acktr/utils.py:8-35 (variant 2) / Space.check_box, envs/bpp0/space.py:111-144 (variant 1)."""
    return int(check_box_batch(np.asarray(plain)[None], [[x, y, lx, ly, z]], container_size, variant)[0])


def check_box_batch(plains, queries, container_size, variant=_cabi.MASK_V2):
    """plains [n, L, W] (or [n, LW]), queries [n, 5] = (x, y, lx, ly, z) -> np.int32[n]."""
    L, W, H = (int(v) for v in container_size)
    dev = torch.device("cuda", torch.cuda.current_device())
    p = torch.as_tensor(np.asarray(plains)).to(device=dev, dtype=torch.int32).reshape(-1, L * W).contiguous()
    q = torch.as_tensor(np.asarray(queries)).to(device=dev, dtype=torch.int32).reshape(-1, 5).contiguous()
    assert p.shape[0] == q.shape[0]
    out = torch.empty((p.shape[0],), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().bpp_check_box(C.c_void_p(p.data_ptr()), C.c_void_p(q.data_ptr()), p.shape[0], L, W, H,
                                              int(variant), C.c_void_p(out.data_ptr()), _stream(dev)))
    return out.cpu().numpy()


def mask_from_plain(plains, boxes, container_size, enable_rotation=False, variant=_cabi.MASK_V2):
    """plains int [n, LW], boxes int [n, 3] -> float32 CUDA tensor [n, LW(1+rot)]."""
    L, W, H = (int(v) for v in container_size)
    dev = torch.device("cuda", torch.cuda.current_device())
    p = torch.as_tensor(np.asarray(plains)).to(device=dev, dtype=torch.int32).reshape(-1, L * W).contiguous()
    b = torch.as_tensor(np.asarray(boxes)).to(device=dev, dtype=torch.int32).reshape(-1, 3).contiguous()
    out = torch.empty((p.shape[0], L * W * (2 if enable_rotation else 1)), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().bpp_mask_from_plain(C.c_void_p(p.data_ptr()), C.c_void_p(b.data_ptr()), p.shape[0], L, W,
                                                    H, int(bool(enable_rotation)), int(variant),
                                                    C.c_void_p(out.data_ptr()), _stream(dev)))
    return out


def get_vec_normalize(venv):
    """acktr/utils.py:96-102: the batched env has no VecNormalize in its chain."""
    return None
