#!/bin/bash
# Round-2 session V (1 GPU): the default bench line as the driver runs it (all legs), smoke, and the ncu evidence
# of the same command at HEAD (launch list, DRAM traffic of the two dominant kernels)
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v_smoke.log 2>&1; tail -1 gpurun_out/v_smoke.log
( time timeout 1200 python bench.py > gpurun_out/v_bench_n1.log 2> gpurun_out/v_bench_n1.err ) 2>&1 | tail -3
python - <<PY
import json
for line in open("gpurun_out/v_bench_n1.log"):
    if line.startswith("{"):
        d = json.loads(line)
        print("N=1: %.1f it/s  %.3f ms/step  parity %s  orth frac %.4f  e2e %s  mgs %s cpu %s" % (
            d["value"], d["ms_per_step"], d.get("parity_vs_cpu_max_rel"), d["roofline"]["frac"],
            (d.get("e2e") or {}).get("value"), (d.get("mgs_value") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value")))
        for k, v in (d.get("configs") or {}).items():
            print("   ", k, v.get("it_per_s"), v.get("frac_of_measured_peak"), (v.get("parity_vs_reference") or {}).get("max_rel_updated"), v.get("error"))
PY
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs --no-mgs"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 220 --csv --log-file gpurun_out/v_launch_list.csv $B > gpurun_out/v_ncu_a.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'orth_kernel|spmv_staged' -s 186 -c 62 --csv --log-file gpurun_out/v_traffic.csv $B > gpurun_out/v_ncu_b.log 2>&1
tail -2 gpurun_out/v_ncu_b.log
