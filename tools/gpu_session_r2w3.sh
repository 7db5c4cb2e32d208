#!/bin/bash
# Round-2 session W3 (1 GPU, the last GPU seconds of the round): the bench line with the L2 residency window on
# (value, e2e, mgs_value, C5) and off on the same box, a half-window variant, the full-size GPU tests with the window on
set -u
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
( time KRY_L2_WINDOW=1 timeout 40 $B --extra c5 > gpurun_out/w3_on.log 2> gpurun_out/w3_on.err ) 2>&1 | grep real
( time KRY_L2_WINDOW=0 timeout 40 $B --extra c5 > gpurun_out/w3_off.log 2> gpurun_out/w3_off.err ) 2>&1 | grep real
( time KRY_L2_WINDOW=1 KRY_L2_WINDOW_RATIO=0.5 timeout 20 $B --no-e2e --no-extra-configs > gpurun_out/w3_half.log 2> gpurun_out/w3_half.err ) 2>&1 | grep real
python - <<PY
import json
for tag in ("on", "off", "half"):
    for line in open("gpurun_out/w3_%s.log" % tag):
        if line.startswith("{"):
            d = json.loads(line)
            c5 = (d.get("configs") or {}).get("c5", {})
            print("%-4s %.1f it/s orth %.1f us spmv %.1f us e2e %s mgs %s c5 %s final %r l2 %s" % (
                tag, d["value"], 1e3 * d["roofline"]["avg_launch_ms"], 1e3 * d["roofline_spmv"]["avg_launch_ms"],
                (d.get("e2e") or {}).get("value"), (d.get("mgs_value") or {}).get("value"), c5.get("it_per_s"),
                d["final_resnorm"], d.get("l2_window")))
PY
( time KRY_L2_WINDOW=1 timeout 30 python -m pytest tests/test_fullsize_gpu.py -q -x -p no:cacheprovider > gpurun_out/w3_fullsize.log 2>&1 ) 2>&1 | grep real
tail -2 gpurun_out/w3_fullsize.log
