#!/bin/bash
set -u
mkdir -p gpurun_out
for c in c5 c3 c2; do
  timeout 300 python tools/profile_solver.py $c > gpurun_out/e_prof_$c.txt 2>&1; cat gpurun_out/e_prof_$c.txt | tail -20
done
timeout 300 python tools/profile_solver.py c5 --host > gpurun_out/e_prof_c5_host.txt 2>&1; head -50 gpurun_out/e_prof_c5_host.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/e_bench.log 2>&1; tail -1 gpurun_out/e_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
