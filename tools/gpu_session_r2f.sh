#!/bin/bash
# Round-2 session F (2 GPUs): dist tests + C2 bench at N=2 with halo-from-q on and off
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/f_pytest_dist.log 2>&1
echo "pytest exit $?" >> gpurun_out/f_pytest_dist.log; tail -15 gpurun_out/f_pytest_dist.log
for hq in 1 0; do
KRY_DIST_HALO_FROM_Q=$hq timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/f_bench_n2_hq$hq.log 2>&1
tail -1 gpurun_out/f_bench_n2_hq$hq.log | cut -c1-220
done
KRY_TRACE=1 timeout 300 python tools/e2e_trace.py > gpurun_out/f_e2e_trace.txt 2>&1; tail -12 gpurun_out/f_e2e_trace.txt
KRY_TRACE=1 timeout 300 python tools/e2e_trace.py nogc > gpurun_out/f_e2e_trace_nogc.txt 2>&1; tail -12 gpurun_out/f_e2e_trace_nogc.txt
