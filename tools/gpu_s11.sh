#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "int32 or packed or host_step or golden_rollout or vecenv or headline or config or pageable" > gpurun_out/pytest_s11.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_s11.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-aux --no-cpu-baseline > gpurun_out/bench_s11.json 2> gpurun_out/bench_s11.err; echo "bench exit $?"; tail -2 gpurun_out/bench_s11.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_s11.json') if l.startswith('{')][0]); print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'e2e_py', d['e2e_py']['ms_per_step'])"
