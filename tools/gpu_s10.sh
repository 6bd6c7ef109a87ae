#!/bin/bash
# small-env batch rule: auto vs forced batch sizes, single-step and 16 steps per launch
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
for bpw in auto 0 1 2 3 4; do
if [ "$bpw" = auto ]; then unset BPP_BPW_LOG2; else export BPP_BPW_LOG2=$bpw; fi
python - <<'PY'
import json, sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'online-3d-bpp-drl_b200')
import torch, bench, b200bpp
from b200bpp import _cabi
dev = torch.device('cuda', 0)
out = {}
for name, n in (('cfg1', 1024), ('cfg1', 4096), ('cfg1', 8192), ('cfg1', 16384), ('cfg1', 32768), ('target_rot', 4096), ('target_rot', 16384)):
    r = bench.Runner(b200bpp, _cabi, name, dev, 1234, 0, 1, bins=n)
    single = r.time_device(300, 20)[0]
    multi = bench.time_multi_step(r, K=16, launches=30)['ms_per_env_step'] if True else None
    out['%s_%d' % (name, n)] = [round(single * 1e3, 2), round(multi * 1e3, 2)]
    r.close()
print('bpw', os.environ.get('BPP_BPW_LOG2', 'auto'), json.dumps(out))
PY
done > gpurun_out/s10_small_envs.txt 2>&1
cat gpurun_out/s10_small_envs.txt
