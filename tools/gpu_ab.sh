#!/bin/bash
# One bounded GPU session (gpurun): time the kernel build variants against the default build, then run the GPU parity
# suite on the fastest one.  Everything lands in gpurun_out/ as it is produced.
#   python online-3d-bpp-drl_b200/build.py --variants        # HERE, before the call: nvcc cross-compiles without a GPU
#   /usr/local/graft/bin/gpurun --timeout 420 -- 'AB_NCU=1 AB_SANITIZE=1 bash tools/gpu_ab.sh'
# A session costs ~2.5 GPU-minutes of fixed overhead plus: A/B ~15 s, GPU suite ~35 s, ncu ~45 s, sanitizer ~2 min.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
V=online-3d-bpp-drl_b200/b200bpp
LIBS="default=$V/libbpp_b200.so"
for n in ${AB_VARIANTS:-baseline split nolut words lsu tier all occ5}; do
    [ -f "$V/variants/libbpp_b200_$n.so" ] && LIBS="$LIBS $n=$V/variants/libbpp_b200_$n.so"
done
# compute-sanitizer over every kernel family of the default build (AB_SANITIZE=1; ~1 min per tool)
if [ -n "$AB_SANITIZE" ]; then
    for tool in memcheck racecheck; do
        timeout 150 compute-sanitizer --tool $tool python tools/sanitize_smoke.py > gpurun_out/sanitizer_$tool.log 2>&1
        echo "$tool exit $?" >> gpurun_out/sanitizer_$tool.log
        tail -2 gpurun_out/sanitizer_$tool.log
    done
fi
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/ab_gpu.txt 2>&1
timeout "${AB_TIMEOUT:-170}" python tools/ab_bench.py --libs $LIBS --workloads ${AB_WORKLOADS:-target target_rot cfg3 cfg4} \
    --steps "${AB_STEPS:-200}" --rounds "${AB_ROUNDS:-3}" --bw > gpurun_out/ab.jsonl 2> gpurun_out/ab.err
echo "ab_bench exit $?" >> gpurun_out/ab.err
BEST=$(python tools/ab_pick.py gpurun_out/ab.jsonl)
echo "fastest (geometric mean over the workloads): $BEST" | tee gpurun_out/ab_best.txt
if [ "$BEST" = "default" ]; then LIB=$PWD/$V/libbpp_b200.so; else LIB=$PWD/$V/variants/libbpp_b200_$BEST.so; fi
BPP_B200_LIB=$LIB timeout "${PYTEST_TIMEOUT:-300}" python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$BEST.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_$BEST.log
tail -4 gpurun_out/pytest_$BEST.log
cat gpurun_out/ab.jsonl
tail -3 gpurun_out/ab.err
# one ncu --set full capture of a steady-state step launch of the chosen build (profiles/summarize.py reads it here)
if [ -n "$AB_NCU" ]; then
    BPP_B200_LIB=$LIB timeout "${NCU_TIMEOUT:-75}" ncu --set full --clock-control none --import-source on \
        -k regex:bpp_step_kernel --launch-skip 6 --launch-count 1 -f -o gpurun_out/prof_$BEST \
        python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 10 > gpurun_out/ncu_$BEST.log 2>&1
    echo "ncu exit $?" >> gpurun_out/ncu_$BEST.log
    tail -2 gpurun_out/ncu_$BEST.log
fi
