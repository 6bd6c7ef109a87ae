#!/bin/bash
# GPU session 2 of round 2: TMA probe, GPU suite on the rebuilt library, generator timing + ncu capture.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 120 tools/_build/tma_probe > gpurun_out/tma_probe.jsonl 2>&1; echo "probe exit $?"; cat gpurun_out/tma_probe.jsonl
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s2.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_s2.log
for w in cut1 cut2; do
  timeout 120 python bench.py --workload $w --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 10 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w exit $?"; cut -c1-400 gpurun_out/bench_$w.json
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:bpp_gen_kernel --launch-skip 12 --launch-count 1 -f -o gpurun_out/prof_gen_$w \
      python bench.py --workload $w --steps 5 --warmup 10 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_gen_$w.log 2>&1; echo "ncu $w exit $?"
done
