#!/bin/bash
# Full validation of the in-tree build on one B200: GPU suite, bench, sanitizer, ncu captures, launch list.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
T=${1:-fin}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_$T.log
timeout 500 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench exit $?"; tail -2 gpurun_out/bench_$T.err
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err; echo "ref arm exit $?"
for tool in memcheck racecheck; do
  for f in 3 4 5; do
    BPP_TMA_FORCE=$f timeout 240 compute-sanitizer --tool $tool python tools/sanitize_smoke.py > gpurun_out/sanitizer_${tool}_tma${f}_$T.log 2>&1
    echo "$tool tma$f exit $?"; tail -1 gpurun_out/sanitizer_${tool}_tma${f}_$T.log
  done
done
for w in target target_rot cfg2 cfg3; do
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:bpp_step_kernel --launch-skip 14 --launch-count 1 -f -o gpurun_out/prof_${T}_$w \
      python bench.py --workload $w --steps 5 --warmup 10 --no-cpu-baseline --no-aux --e2e-steps 3 > gpurun_out/ncu_${T}_$w.log 2>&1; echo "ncu $w exit $?"
done
for w in cut1 cut2; do
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:bpp_gen --launch-skip 12 --launch-count 1 -f -o gpurun_out/prof_${T}_gen_$w \
      python bench.py --workload $w --steps 5 --warmup 10 --no-cpu-baseline --no-aux --e2e-steps 3 > gpurun_out/ncu_${T}_gen_$w.log 2>&1; echo "ncu gen $w exit $?"
done
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/launches_$T.log 2>&1; echo "launch list exit $?"
