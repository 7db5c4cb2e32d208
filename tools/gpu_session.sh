#!/bin/bash
# One GPU-box session (gpurun -- 'bash tools/gpu_session.sh'): the GPU test tier, the default bench
# line, the ncu launch list of the same command and the full-size configurations, with everything
# worth keeping written under gpurun_out/ (copied to profiles/ by hand afterwards).
# First thing to run in a round: it covers the tests that were written without GPU time
# (tests/test_zcomplex_gpu.py, tests/test_zz_analysis_gpu.py).
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
# the files that sort last, again without -x, so one failure does not hide the rest
timeout 900 python -m pytest tests/test_zcomplex_gpu.py tests/test_zz_analysis_gpu.py -m gpu -q -p no:cacheprovider \
    > gpurun_out/pytest_new.log 2>&1
tail -15 gpurun_out/pytest_new.log
# opt-in variants (new cooperative kernel): own process, own timeout
KRY_TEST_VARIANTS=1 timeout 600 python -m pytest tests/test_zz_variants_gpu.py -m gpu -q -p no:cacheprovider \
    > gpurun_out/pytest_variants.log 2>&1
tail -5 gpurun_out/pytest_variants.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
gcc -O2 -I include -I /usr/local/cuda/include examples/gmres_c_abi.c -o gpurun_out/gmres_c_abi -L krypy_b200 -lkrypy_b200 \
    -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/krypy_b200 && timeout 300 gpurun_out/gmres_c_abi 1024 30 5 \
    > gpurun_out/c_abi_example.log 2>&1; tail -2 gpurun_out/c_abi_example.log
timeout 900 python bench.py > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log
# A/B of the two opt-in measurement switches (correctness first, then the bench line without the CPU legs)
for sw in KRY_ORTH_SMALLK KRY_ORTH_SPLIT_SCALE KRY_ORTH_CUNROLL; do
  env $sw=1 timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_solvers_gpu.py -m gpu -q -x -p no:cacheprovider \
      -k "orth or fixture or arnoldi or variants" > gpurun_out/pytest_$sw.log 2>&1; tail -1 gpurun_out/pytest_$sw.log
  env $sw=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_$sw.log 2>&1; tail -1 gpurun_out/bench_$sw.log | cut -c1-200
done
# the small-tile variant across many tiles (every CGS call up to 40 vectors goes through it)
KRY_ORTH_SMALLK=40 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "orth" \
    > gpurun_out/pytest_smallk40.log 2>&1; tail -1 gpurun_out/pytest_smallk40.log
for thr in 2 8 12; do
  KRY_ORTH_SMALLK=$thr timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_smallk_$thr.log 2>&1
  tail -1 gpurun_out/bench_smallk_$thr.log | cut -c1-160
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 220 --csv \
    --log-file gpurun_out/launch_list.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e \
    > gpurun_out/ncu_bench.log 2>&1
# full ncu capture of the fused Gram-Schmidt kernel at k = 0, 1, 2 (launches 60..62 = start of the 3rd cycle)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:orth_kernel -s 60 -c 3 -f \
    -o gpurun_out/orth_smallk python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_orth_smallk.log 2>&1
ncu -i gpurun_out/orth_smallk.ncu-rep --page details > gpurun_out/orth_smallk_details.txt 2>/dev/null
timeout 1200 python tools/run_configs.py c2 c3 c4 c4r c5 > gpurun_out/configs_fullsize.json 2> gpurun_out/configs.err
tail -3 gpurun_out/configs.err
# C5 with the fused diagonal-ip_B Lanczos kernel (opt-in) next to the default above
KRY_LANCZOS_DIAGB=1 timeout 900 python tools/run_configs.py c5 > gpurun_out/configs_c5_lanczos_diagB.json 2>> gpurun_out/configs.err
# optional (slow, ~10-50x): memory / race checks of the kernels on the small parity cases
#   compute-sanitizer --tool memcheck python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "not 100003" > gpurun_out/memcheck.log 2>&1
#   compute-sanitizer --tool racecheck python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "block_dot or orth_fused" > gpurun_out/racecheck.log 2>&1
