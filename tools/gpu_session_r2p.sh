#!/bin/bash
# Round-2 session P (1 GPU): per-kernel times of the one-wait Arnoldi step at the per-rank size of an 8-GPU
# C2 run, emulated on one GPU (tools/bench_dist_kernels.py)
set -u
mkdir -p gpurun_out
timeout 600 python tools/bench_dist_kernels.py 3162 ${KS:-0,4,8,11,12,15,20,25,29} > gpurun_out/p_dist_kernels_8.txt 2>&1; grep -v "^\[\|NCCL\|Warn" gpurun_out/p_dist_kernels_8.txt | tail -34
