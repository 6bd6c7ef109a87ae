#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
V=online-3d-bpp-drl_b200/b200bpp
timeout 250 python tools/ab_bench.py --libs default=$V/libbpp_b200.so t20_3=$V/variants/libbpp_b200_t20_3.so t20_4=$V/variants/libbpp_b200_t20_4.so \
   --workloads cfg3 --steps 200 --rounds 3 > gpurun_out/ab_s15.jsonl 2> gpurun_out/ab_s15.err; echo "ab exit $?"; tail -2 gpurun_out/ab_s15.err
