#!/bin/bash
# Round-2 session G: full default bench at N=1 (extras included) and the N=2 line with e2e/parity/extras
set -u
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  ( time timeout 850 python bench.py > gpurun_out/g_bench_n1.log 2> gpurun_out/g_bench_n1.err ) 2>&1 | tail -3
  tail -12 gpurun_out/g_bench_n1.err
  tail -1 gpurun_out/g_bench_n1.log | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['step_ms_wall'],'mgs',d.get('mgs_value'),'parity',d.get('parity_vs_cpu_max_rel'))
print('roofline',d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'])
for k,v in d.get('configs',{}).items(): print(k, {a:b for a,b in v.items() if a not in ('config','first4','U')})
"
else
  ( time timeout 850 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus $N --steps 10 --warmup 3 ${2:-} > gpurun_out/g_bench_n$N.log 2> gpurun_out/g_bench_n$N.err ) 2>&1 | tail -3
  grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/g_bench_n$N.err | tail -12
  tail -1 gpurun_out/g_bench_n$N.log | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print('value',d['value'],'e2e',d['e2e'],'mgs',d.get('mgs_value'),'parity',d.get('parity_vs_cpu_max_rel'))
for k,v in d.get('configs',{}).items(): print(k, {a:b for a,b in v.items() if a not in ('config','first4','U')})
"
fi
