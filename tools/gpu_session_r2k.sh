#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_solvers_gpu.py -m gpu -x -q -p no:cacheprovider -k "gram or cholqr or cg_device or deflat or defl" > gpurun_out/k_pytest.log 2>&1; tail -4 gpurun_out/k_pytest.log
timeout 300 python tools/profile_solver.py c4 --maxiter 0 > gpurun_out/k_prof_c4_setup.txt 2>&1; tail -14 gpurun_out/k_prof_c4_setup.txt
timeout 300 python tools/profile_solver.py c4 --maxiter 60 > gpurun_out/k_prof_c4.txt 2>&1; tail -14 gpurun_out/k_prof_c4.txt
