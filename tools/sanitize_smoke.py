"""Small workload for compute-sanitizer (memcheck / racecheck): every kernel family once, few bins, few steps.

    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "online-3d-bpp-drl_b200"))

import numpy as np   # noqa: E402
import torch         # noqa: E402
import b200bpp       # noqa: E402
from b200bpp import utils   # noqa: E402

import ctypes as C      # noqa: E402
from b200bpp import _cabi   # noqa: E402

for kw in (dict(data_type="rs"), dict(data_type="rs", enable_rotation=True), dict(data_type="cut1", enable_rotation=True),
           dict(data_type="cut2"), dict(data_type="rs", container_size=(20, 20, 20)),
           dict(data_type="rs", container_size=(7, 9, 6), box_set=[(1, 2, 1), (3, 3, 2), (8, 2, 1)])):
    env = b200bpp.BinPackingVecEnv(96, seed=5, want_sample=True, **kw)
    env.reset()
    for t in range(8):
        env.step_device(env.sample_t.clone())
    env.step(env.sample_t.view(-1, 1))           # host entry point (zero-copy records, device actions)
    env.step(env.sample_t.cpu().view(-1, 1))     # host entry point (pinned host actions)
    if kw["data_type"] == "rs":                  # packed records, four env steps in one launch
        res = torch.zeros((96, 4), dtype=torch.int32, device="cuda")
        _cabi.check(_cabi.lib().bpp_step_packed(env._h, C.c_void_p(env.sample_t.data_ptr()), C.c_void_p(env.obs.data_ptr()),
                                                C.c_void_p(env.location_masks.data_ptr()), C.c_void_p(res.data_ptr()),
                                                C.c_void_p(env.done_t.data_ptr()), C.c_void_p(env.sample_t.data_ptr()), 4,
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
    env.reset()                                  # explicit second reset
    env.step_device(env.sample_t.clone())
    sd = env.state_dict()
    env.load_state_dict(sd)
    print(kw, env.metrics())
    env.close()
obs = torch.zeros(5, 400, device="cuda")
obs[:, 100:] = 3
print(utils.mask_batch(obs, (10, 10, 10), True).sum().item(),
      utils.check_box_batch(np.zeros((4, 10, 10)), [[2, 2, 0, 0, 2]] * 4, (10, 10, 10)).tolist())
print("sanitize smoke done")
