#!/bin/bash
# Round-2 session O (N GPUs): the one-wait Arnoldi step on real NVLink peers: parity test (N = 2), then
# A/B of the C2 bench: KRY_DIST_FUSED=1 (SpMV, dot+<w,w>, update+scale+halo+Givens: one cross-GPU wait)
# against KRY_DIST_FUSED=0 (dot / update / scale+halo / Givens: two waits)
set -u
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_dist_gpu.py tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "dist or partitioned or mdot" > gpurun_out/o_pytest_n$N.log 2>&1
  echo "pytest exit $?" >> gpurun_out/o_pytest_n$N.log; tail -5 gpurun_out/o_pytest_n$N.log
fi
B="bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs --no-mgs"
for f in ${AB:-1 0 1 0}; do
  KRY_DIST_FUSED=$f timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2967$f \
      $B > gpurun_out/o_bench_n${N}_fused$f.log 2> gpurun_out/o_bench_n${N}_fused$f.err
  python - <<PY
import json
for line in open("gpurun_out/o_bench_n${N}_fused$f.log"):
    if line.startswith("{"):
        d = json.loads(line)
        print("N=$N fused=$f: %.1f it/s  %.3f ms/step  final %.16e  parity %s  launches %s" % (d["value"], d["ms_per_step"], d["final_resnorm"], d.get("parity_vs_cpu_max_rel"), d.get("gpu_launches")))
        print("   orth_by_nv", {k: v["us"] for k, v in list(d.get("orth_by_nv", {}).items())[:60]})
PY
done
