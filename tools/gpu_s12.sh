#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 200 python tools/ab_bench.py --libs default=online-3d-bpp-drl_b200/b200bpp/libbpp_b200.so prev=online-3d-bpp-drl_b200/b200bpp/variants/libbpp_b200_prev.so \
   --workloads target target_rot cfg2 cfg3 cfg4 --steps 200 --rounds 3 > gpurun_out/ab_s12.jsonl 2> gpurun_out/ab_s12.err; echo "ab exit $?"
timeout 600 python -m pytest tests -m gpu -x -q -k "headline or golden or config or odd or extreme or cut or packed or policy or lookahead" > gpurun_out/pytest_s12.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_s12.log
