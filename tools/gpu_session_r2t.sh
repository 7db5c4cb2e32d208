#!/bin/bash
# Round-2 session T (1 GPU): pre-launched restart cycles: graph-replay test, 2/3-rank parity, cycle-boundary profile
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_solvers_gpu.py tests/test_dist_gpu.py -m gpu -x -q -p no:cacheprovider -k "graph_replay or partitioned" > gpurun_out/t_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/t_pytest.log; tail -5 gpurun_out/t_pytest.log
timeout 600 python tools/cycle_boundary_profile.py 1118 > gpurun_out/t_cycle_boundary.txt 2>&1
grep -v "OMP_NUM\|^\*\*\*\|NCCL version\|^\[W" gpurun_out/t_cycle_boundary.txt | head -${LINES_MAX:-48}
