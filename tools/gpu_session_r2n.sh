#!/bin/bash
# Round-2 session N (1 GPU): the one-wait distributed Arnoldi step (kry_spmv_csr_mdot + kry_dist_update_scale):
# kernel test, 2- and 3-rank parity (ranks share the GPU), kernel timing of the multi-dot SpMV
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "mdot or spmv" > gpurun_out/n_pytest_mdot.log 2>&1
echo "pytest exit $?" >> gpurun_out/n_pytest_mdot.log; tail -5 gpurun_out/n_pytest_mdot.log
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/n_pytest_dist.log 2>&1
echo "pytest exit $?" >> gpurun_out/n_pytest_dist.log; tail -30 gpurun_out/n_pytest_dist.log
timeout 300 python tools/bench_mdot.py > gpurun_out/n_bench_mdot.txt 2>&1; cat gpurun_out/n_bench_mdot.txt
