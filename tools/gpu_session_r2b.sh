#!/bin/bash
# Round-2 session B (N GPUs): named multi-GPU configs through tools/run_configs_dist.py
set -u
N=${1:-2}; shift
mkdir -p gpurun_out
for c in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 \
      tools/run_configs_dist.py $c > gpurun_out/b_${c}_${N}gpu.json 2> gpurun_out/b_${c}_${N}gpu.err
  echo "$c on $N GPUs: exit $?"; tail -c 1800 gpurun_out/b_${c}_${N}gpu.json; tail -5 gpurun_out/b_${c}_${N}gpu.err
done
