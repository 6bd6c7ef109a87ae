#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "cut or identity or generator or checkpoint or packed or host_step" > gpurun_out/pytest_s7.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_s7.log
for w in cut1 cut2; do
  timeout 120 python bench.py --workload $w --steps 200 --warmup 20 --no-cpu-baseline --no-aux --e2e-steps 10 > gpurun_out/bench_s7_$w.json 2> gpurun_out/bench_s7_$w.err; echo "bench $w exit $?"; python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_s7_$w.json') if l.startswith('{')][0]); print('$w', d['ms_per_step'], d['value'])"
done
[ "$(nvidia-smi -L | wc -l)" -ge 2 ] || { echo "one GPU: skipping the 2-rank runs"; exit 0; }
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_s7_n2.json 2> gpurun_out/bench_s7_n2.err; echo "bench n2 exit $?"; tail -2 gpurun_out/bench_s7_n2.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 4 --warmup 1 --ref-seconds 5 --no-port > gpurun_out/bench_s7_ref_n2.json 2> gpurun_out/bench_s7_ref_n2.err; echo "ref n2 exit $?"
