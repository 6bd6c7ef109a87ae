#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 200 python tools/ab_bench.py --libs default=online-3d-bpp-drl_b200/b200bpp/libbpp_b200.so w7=online-3d-bpp-drl_b200/b200bpp/variants/libbpp_b200_w7.so \
   --workloads target target_rot cfg2 cfg4 --steps 200 --rounds 3 > gpurun_out/ab_s13.jsonl 2> gpurun_out/ab_s13.err; echo "ab exit $?"; tail -2 gpurun_out/ab_s13.err
