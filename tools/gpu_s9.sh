#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
AB_VARIANTS="split words split_words" AB_WORKLOADS="target target_rot cfg2 cfg4" AB_TIMEOUT=200 PYTEST_TIMEOUT=5 bash tools/gpu_ab.sh > gpurun_out/s9.log 2>&1
cat gpurun_out/ab.jsonl | head -c 100 > /dev/null
for bpw in 0 1 2; do
BPP_BPW_LOG2=$bpw python - <<'PY'
import json, sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'online-3d-bpp-drl_b200')
import torch, bench, b200bpp
from b200bpp import _cabi
dev = torch.device('cuda', 0)
r = bench.Runner(b200bpp, _cabi, 'cfg1', dev, 1234, 0, 1)
print('bpw', os.environ['BPP_BPW_LOG2'], 'single', r.time_device(300, 20)[0], 'multi16', json.dumps(bench.time_multi_step(r, K=16, launches=40)))
r.close()
PY
done
