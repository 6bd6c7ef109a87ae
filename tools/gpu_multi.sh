#!/bin/bash
# bench.py under torchrun on the GPUs of this box (both arms), as the driver launches it
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
G=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/topo_$G.txt 2>&1
for n in ${NLIST:-1 2 4 8}; do
  [ $n -le $G ] || break
  if [ $n = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/multi_n$n.json 2> gpurun_out/multi_n$n.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/multi_n$n.json 2> gpurun_out/multi_n$n.err
  fi
  echo "n=$n exit $?"; tail -1 gpurun_out/multi_n$n.err | cut -c1-200
done
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29600 bench.py --impl reference --gpus $G --steps 4 --warmup 1 --ref-seconds 5 --no-port > gpurun_out/multi_ref.json 2> gpurun_out/multi_ref.err; echo "ref exit $?"
