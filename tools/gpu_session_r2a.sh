#!/bin/bash
# Round-2 session A: everything that never ran on a B200, without -x, plus the recycling diagnostic.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > gpurun_out/a_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/a_pytest_gpu.log
tail -30 gpurun_out/a_pytest_gpu.log
KRY_TEST_VARIANTS=1 timeout 600 python -m pytest tests/test_zz_variants_gpu.py -m gpu -q -p no:cacheprovider \
    > gpurun_out/a_pytest_variants.log 2>&1
tail -8 gpurun_out/a_pytest_variants.log
timeout 300 python tools/diag_recycling.py gpu gpurun_out/diag_recycling_gpu.json > gpurun_out/a_diag.log 2>&1; tail -21 gpurun_out/a_diag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; tail -2 gpurun_out/a_smoke.log
gcc -O2 -I include -I /usr/local/cuda/include examples/gmres_c_abi.c -o gpurun_out/gmres_c_abi -L krypy_b200 -lkrypy_b200 \
    -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/krypy_b200 && timeout 300 gpurun_out/gmres_c_abi 1024 30 5 \
    > gpurun_out/a_c_abi_example.log 2>&1; tail -2 gpurun_out/a_c_abi_example.log
timeout 600 python bench.py > gpurun_out/a_bench_n1.log 2>&1; tail -1 gpurun_out/a_bench_n1.log | cut -c1-600
for sw in KRY_ORTH_SMALLK KRY_ORTH_SPLIT_SCALE KRY_ORTH_CUNROLL; do
  env $sw=1 timeout 400 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/a_bench_$sw.log 2>&1; tail -1 gpurun_out/a_bench_$sw.log | cut -c1-200
done
for thr in 2 8; do
  KRY_ORTH_SMALLK=$thr timeout 400 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/a_bench_smallk_$thr.log 2>&1
  tail -1 gpurun_out/a_bench_smallk_$thr.log | cut -c1-160
done
timeout 900 python tools/run_configs.py c4 c4r c5 > gpurun_out/a_configs.json 2> gpurun_out/a_configs.err
tail -3 gpurun_out/a_configs.err
KRY_LANCZOS_DIAGB=1 timeout 600 python tools/run_configs.py c5 > gpurun_out/a_configs_c5_diagB.json 2>> gpurun_out/a_configs.err
tail -c 1500 gpurun_out/a_configs.json; tail -c 800 gpurun_out/a_configs_c5_diagB.json
