#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
for c in 9 4 3; do
BPP_MAX_CTAS=$c python - <<'PY'
import json, sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'online-3d-bpp-drl_b200')
import torch, bench, b200bpp
from b200bpp import _cabi
dev = torch.device('cuda', 0)
out = {}
for name in ('target', 'target_rot', 'cfg2', 'cfg3', 'cfg4'):
    r = bench.Runner(b200bpp, _cabi, name, dev, 1234, 0, 1)
    best = min(r.time_device(150, 10)[0] for _ in range(3))
    out[name] = round(best, 4)
    r.close()
print('max_ctas', os.environ['BPP_MAX_CTAS'], json.dumps(out))
PY
done > gpurun_out/s16_ctas.txt 2>&1
cat gpurun_out/s16_ctas.txt
