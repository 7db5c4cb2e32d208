#!/bin/bash
# Round-2 session W2 (1 GPU): A/B of the L2 residency window on w = A v_k (tools/bench_l2window.py: C2 cgs / mgs and
# C5, window off / on, default and side stream, bitwise identity of the histories), then one bench line with the window on
set -u
mkdir -p gpurun_out
( time timeout 60 python tools/bench_l2window.py > gpurun_out/w2_l2window.json 2> gpurun_out/w2_l2window.err; echo "exit $?" >> gpurun_out/w2_l2window.err ) 2>&1 | grep real
tail -14 gpurun_out/w2_l2window.err
( time KRY_L2_WINDOW=1 timeout 45 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs > gpurun_out/w2_bench_window1.log 2> gpurun_out/w2_bench_window1.err; echo "exit $?" >> gpurun_out/w2_bench_window1.err ) 2>&1 | grep real
python - <<PY
import json
for line in open("gpurun_out/w2_bench_window1.log"):
    if line.startswith("{"):
        d = json.loads(line)
        print("window1: %.1f it/s  orth frac %.4f  mgs %s  l2 %s parity %s" % (d["value"], d["roofline"]["frac"],
              (d.get("mgs_value") or {}).get("value"), d.get("l2_window"), d.get("parity_vs_cpu_max_rel")))
PY
