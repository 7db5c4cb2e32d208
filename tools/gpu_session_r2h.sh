#!/bin/bash
# Round-2 session H (N GPUs): per-kernel device time of the row-partitioned C3 / C5 iterations (rank 0)
set -u
N=${1:-4}
mkdir -p gpurun_out
for c in "c3 --grid 400" "c5 --grid 4000"; do
  set -- $c
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29661 \
      tools/profile_solver.py $c --maxiter 50 > gpurun_out/h_prof_$1_n$N.txt 2>&1
  grep -v "OMP_NUM\|^\*\*\*\|^$\|NCCL version" gpurun_out/h_prof_$1_n$N.txt | tail -16
done
