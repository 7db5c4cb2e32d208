#!/bin/bash
# Round-2 session I (2 GPUs): whole GPU tier (-x) + C3/C5 at 2 GPUs with the pipelined CG
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/i_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/i_pytest_gpu.log; tail -8 gpurun_out/i_pytest_gpu.log
timeout 600 python tools/run_configs_parity.py c3 --no-parity > gpurun_out/i_c3_1gpu.json 2> gpurun_out/i_c3_1gpu.err; grep "c3:" gpurun_out/i_c3_1gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 \
   tools/run_configs_parity.py c3 c5 --steps-c3 6 --steps-c5 6 > gpurun_out/i_c3c5_2gpu.json 2> gpurun_out/i_c3c5_2gpu.err
grep "c3:\|c5:\|Error\|error" gpurun_out/i_c3c5_2gpu.err | tail
