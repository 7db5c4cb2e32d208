#!/bin/bash
# Round-2 session X (1 GPU): ncu --set full of the two Gram-Schmidt kernels of the one-wait distributed step at the
# per-rank size of an 8-GPU C2 run (world = 1 emulation, tools/bench_dist_kernels.py, k = 15)
set -u
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'dist_update_scale_kernel|dist_dot_kernel' -s 2 -c 4 -f \
    -o gpurun_out/x_dist python tools/bench_dist_kernels.py 3162 15 > gpurun_out/x_ncu.log 2>&1
tail -3 gpurun_out/x_ncu.log
ncu -i gpurun_out/x_dist.ncu-rep --page details > gpurun_out/x_dist_details.txt 2>/dev/null
grep -n "dist_dot_kernel\|dist_update_scale_kernel\|Duration\|DRAM Throughput\|Memory Throughput\|Registers Per\|Achieved Occupancy\|Theoretical Occupancy" gpurun_out/x_dist_details.txt | head -40
