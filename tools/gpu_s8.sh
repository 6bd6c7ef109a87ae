#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
for w in cut1 cut2; do
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:bpp_gen --launch-skip 12 --launch-count 1 -f -o gpurun_out/prof_s8_gen_$w \
      python bench.py --workload $w --steps 5 --warmup 10 --no-cpu-baseline --no-aux --e2e-steps 3 > gpurun_out/ncu_s8_$w.log 2>&1; echo "ncu $w exit $?"
done
timeout 200 python bench.py --workload cfg1 --steps 300 --warmup 20 --no-cpu-baseline --e2e-steps 10 > gpurun_out/bench_s8_cfg1.json 2> gpurun_out/bench_s8_cfg1.err; echo "bench cfg1 exit $?"
python - <<'PY'
import json, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'online-3d-bpp-drl_b200')
import torch, bench, b200bpp
from b200bpp import _cabi
dev = torch.device('cuda', 0)
for K in (4, 16, 64):
    r = bench.Runner(b200bpp, _cabi, 'cfg1', dev, 1234, 0, 1)
    print('multi_step', K, json.dumps(bench.time_multi_step(r, K=K, launches=40)))
    r.close()
PY
