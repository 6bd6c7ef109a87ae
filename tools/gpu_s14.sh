#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q -k "graph" > gpurun_out/pytest_s14.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_s14.log
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_s14.json 2> gpurun_out/bench_s14.err; echo "bench exit $?"; tail -2 gpurun_out/bench_s14.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_s14.json') if l.startswith('{')][0]); print(json.dumps(d['aux']['configs[1]'])[:900])"
