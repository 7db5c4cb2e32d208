#!/bin/bash
# Round-2 session S (1 GPU): GPU idle gaps at the restart-cycle boundary (tools/cycle_boundary_profile.py)
set -u
mkdir -p gpurun_out
timeout 600 python tools/cycle_boundary_profile.py 1118 > gpurun_out/s_cycle_boundary.txt 2>&1
grep -v "OMP_NUM\|^\*\*\*\|NCCL version\|^\[W" gpurun_out/s_cycle_boundary.txt | head -120
