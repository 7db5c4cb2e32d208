#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/j_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/j_pytest_gpu.log; tail -8 gpurun_out/j_pytest_gpu.log
( time timeout 850 python bench.py > gpurun_out/j_bench_n1.log 2> gpurun_out/j_bench_n1.err ) 2>&1 | tail -3
tail -12 gpurun_out/j_bench_n1.err
tail -1 gpurun_out/j_bench_n1.log | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print('value',d['value'],'e2e',d['e2e']['value'],'mgs',d.get('mgs_value',{}).get('value'),'parity',d.get('parity_vs_cpu_max_rel'))
for k,v in d.get('configs',{}).items(): print(k, {a:b for a,b in v.items() if a not in ('config','first4','U','parity_vs_reference')})
"
