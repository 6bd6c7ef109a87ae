#!/bin/bash
# GPU session: suite + bench + ncu + launch list + sanitizer on the current default build.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s5.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_s5.log
timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_s5.json 2> gpurun_out/bench_s5.err; echo "bench exit $?"; tail -3 gpurun_out/bench_s5.err
for tool in memcheck racecheck; do
  for f in 4 5; do
    BPP_TMA_FORCE=$f timeout 200 compute-sanitizer --tool $tool python tools/sanitize_smoke.py > gpurun_out/sanitizer_${tool}_tma$f.log 2>&1
    echo "$tool tma$f exit $?"; tail -2 gpurun_out/sanitizer_${tool}_tma$f.log
  done
done
timeout 150 ncu --set full --clock-control none --import-source on -k regex:bpp_step_kernel --launch-skip 14 --launch-count 1 -f -o gpurun_out/prof_s5_target \
    python bench.py --steps 5 --warmup 10 --no-cpu-baseline --no-aux --e2e-steps 3 > gpurun_out/ncu_s5.log 2>&1; echo "ncu exit $?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_s5.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/launches_s5.log 2>&1; echo "launch list exit $?"
