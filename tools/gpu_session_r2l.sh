#!/bin/bash
# Round-2 session L (1 GPU): GPU tier as the driver runs it, C4 set-up profile, ncu evidence (launch list, DRAM
# traffic, --set full of the fused Gram-Schmidt kernel and of the Gram kernel), reference parity at BASELINE 5 truncations
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/l_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/l_pytest_gpu.log; tail -4 gpurun_out/l_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/l_smoke.log 2>&1; tail -1 gpurun_out/l_smoke.log
timeout 300 python tools/profile_solver.py c4 --maxiter 0 > gpurun_out/l_prof_c4_setup.txt 2>&1; tail -9 gpurun_out/l_prof_c4_setup.txt
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra-configs --no-mgs"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 220 --csv --log-file gpurun_out/l_launch_list.csv $B > gpurun_out/l_ncu_a.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'orth_kernel|spmv_staged' -s 62 -c 62 --csv --log-file gpurun_out/l_traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra-configs --no-mgs > gpurun_out/l_ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:orth_kernel -s 75 -c 3 -f -o gpurun_out/l_orth $B > gpurun_out/l_ncu_c.log 2>&1
ncu -i gpurun_out/l_orth.ncu-rep --page details > gpurun_out/l_orth_details.txt 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gram_kernel|block_trsm' -c 4 -f -o gpurun_out/l_gram python tools/profile_solver.py c4 --maxiter 0 > gpurun_out/l_ncu_d.log 2>&1
ncu -i gpurun_out/l_gram.ncu-rep --page details > gpurun_out/l_gram_details.txt 2>/dev/null
ls -la gpurun_out/l_*.ncu-rep
( time timeout 1500 python tools/run_configs_parity.py c5 c4 c3 > gpurun_out/l_configs_parity.json 2> gpurun_out/l_configs_parity.err ) 2>&1 | tail -3
grep "c3:\|c4:\|c5:" gpurun_out/l_configs_parity.err
