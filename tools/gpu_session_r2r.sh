#!/bin/bash
# Round-2 session R (1 GPU): full GPU tier with the whole-cycle graph mode and the one-wait distributed step
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r_pytest_gpu.log; tail -5 gpurun_out/r_pytest_gpu.log
