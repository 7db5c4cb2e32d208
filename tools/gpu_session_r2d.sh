#!/bin/bash
# Round-2 session D (2 GPUs): dist tests + named configs at 2 GPUs
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/d_pytest_dist.log 2>&1
echo "pytest exit $?" >> gpurun_out/d_pytest_dist.log; tail -15 gpurun_out/d_pytest_dist.log
bash tools/gpu_session_r2b.sh 2 c3 c5 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | grep -E "it_per_s|exit|identical|vs_single|Error|error|final_resnorm" 
