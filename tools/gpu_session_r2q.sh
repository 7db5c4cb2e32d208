#!/bin/bash
# Round-2 session Q (N GPUs): whole-cycle CUDA graph against the solver's per-step graphs (tools/probe_cycle_graph.py)
set -u
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29681 \
    tools/probe_cycle_graph.py > gpurun_out/q_probe_n$N.txt 2>&1
grep -v "OMP_NUM\|^\*\*\*\|^$\|NCCL version" gpurun_out/q_probe_n$N.txt | tail -8
