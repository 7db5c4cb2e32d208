#!/bin/bash
# Round-2 session W (N GPUs): the bench line exactly as the driver launches it (all legs) at N > 1
set -u
N=${1:-2}
mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29691 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/w_bench_n$N.log 2> gpurun_out/w_bench_n$N.err ) 2>&1 | tail -3
python - <<PY
import json
ok = False
for line in open("gpurun_out/w_bench_n$N.log"):
    if line.startswith("{"):
        ok = True
        d = json.loads(line)
        print("N=$N: %.1f it/s  %.3f ms/step  parity %s  e2e %s  mgs %s  launches %s" % (
            d["value"], d["ms_per_step"], d.get("parity_vs_cpu_max_rel"), (d.get("e2e") or {}).get("value"),
            (d.get("mgs_value") or {}).get("value"), d.get("gpu_launches")))
        for k, v in (d.get("configs") or {}).items():
            print("   ", k, v.get("it_per_s"), v.get("frac_of_measured_peak"), (v.get("parity_vs_reference") or {}).get("max_rel_updated"), v.get("error"))
if not ok:
    print("NO JSON LINE")
PY
grep -v "OMP_NUM\|^\*\*\*\|^$\|NCCL version" gpurun_out/w_bench_n$N.err | tail -12
