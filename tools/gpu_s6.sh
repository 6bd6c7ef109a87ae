#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s6.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_s6.log
timeout 400 python bench.py --steps 50 --warmup 10 > gpurun_out/bench_s6.json 2> gpurun_out/bench_s6.err; echo "bench exit $?"; tail -3 gpurun_out/bench_s6.err
for w in cut1 cut2; do
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:bpp_gen --launch-skip 12 --launch-count 1 -f -o gpurun_out/prof_s6_gen_$w \
      python bench.py --workload $w --steps 5 --warmup 10 --no-cpu-baseline --no-aux --e2e-steps 3 > gpurun_out/ncu_s6_$w.log 2>&1; echo "ncu $w exit $?"
done
