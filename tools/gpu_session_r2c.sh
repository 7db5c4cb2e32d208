#!/bin/bash
# Round-2 session C: GPU tier as the driver runs it (-x), bench with the rewritten orth kernel, C4r/C5, reference arm
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/c_pytest_gpu.log; tail -6 gpurun_out/c_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/c_bench_n1.log 2>&1; tail -1 gpurun_out/c_bench_n1.log | cut -c1-400
timeout 600 python bench.py --ortho mgs --no-cpu-baseline --no-e2e > gpurun_out/c_bench_mgs.log 2>&1; tail -1 gpurun_out/c_bench_mgs.log | cut -c1-200
timeout 900 python tools/run_configs.py c4 c4r c5 > gpurun_out/c_configs.json 2> gpurun_out/c_configs.err; tail -3 gpurun_out/c_configs.err
python - <<'P'
import json
d=json.load(open('gpurun_out/c_configs.json'))
for k,v in d.items(): print(k, {a:b for a,b in v.items() if a not in ('config','first4')})
P
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c_bench_reference.log 2>&1; tail -1 gpurun_out/c_bench_reference.log | cut -c1-300
[ "${1:-}" = "full" ] && { timeout 900 python tools/reference_full_cycle.py > gpurun_out/c_reference_full_cycle.json 2> gpurun_out/c_reference_full_cycle.err; head -c 600 gpurun_out/c_reference_full_cycle.json | tail -c 300; }
