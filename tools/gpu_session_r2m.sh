#!/bin/bash
# Round-2 session M (1 GPU): GPU tier as the driver runs it (rewritten Gram / block-trsm kernels), C4 set-up and
# iteration profiles (kernel table per iteration)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/m_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/m_pytest_gpu.log; tail -4 gpurun_out/m_pytest_gpu.log
timeout 300 python tools/profile_solver.py c4 --maxiter 0 > gpurun_out/m_prof_c4_setup.txt 2>&1; tail -12 gpurun_out/m_prof_c4_setup.txt
timeout 300 python tools/profile_solver.py c4 --maxiter 60 > gpurun_out/m_prof_c4.txt 2>&1; tail -16 gpurun_out/m_prof_c4.txt
