#!/bin/bash
# Round-2 session U (1 GPU): full GPU tier + the C2 bench line with whole cycles enqueued ahead (eager at N = 1)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/u_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/u_pytest_gpu.log; tail -4 gpurun_out/u_pytest_gpu.log
for a in 1 0; do
KRY_CYCLE_AHEAD=$a timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs --no-mgs > gpurun_out/u_bench_n1_ahead$a.log 2> gpurun_out/u_bench_n1_ahead$a.err
python - <<PY
import json
for line in open("gpurun_out/u_bench_n1_ahead$a.log"):
    if line.startswith("{"):
        d = json.loads(line)
        print("N=1 ahead=$a: %.1f it/s  %.3f ms/step  final %.16e  parity %s  orth frac %.4f spmv frac %.4f  e2e %s" % (
            d["value"], d["ms_per_step"], d["final_resnorm"], d.get("parity_vs_cpu_max_rel"), d["roofline"]["frac"],
            d["roofline_spmv"]["frac"], (d.get("e2e") or {}).get("value")))
PY
done
