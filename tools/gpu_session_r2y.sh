#!/bin/bash
# Round-2 session Y (1 GPU, the last GPU-minutes of the round): the whole GPU tier at HEAD with the native complex
# kernels as the default of the complex path (no -x: every failure is listed), then the complex twin of C2 with the
# native and with the real-embedding kernels (tools/bench_cplx.py)
set -u
mkdir -p gpurun_out
( time timeout 140 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/y_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/y_pytest_gpu.log ) 2>&1 | grep real
tail -5 gpurun_out/y_pytest_gpu.log
grep -E "^(FAILED|ERROR)" gpurun_out/y_pytest_gpu.log | head -20
( time timeout 80 python tools/bench_cplx.py > gpurun_out/y_cplx.json 2> gpurun_out/y_cplx.err; echo "cplx exit $?" >> gpurun_out/y_cplx.err ) 2>&1 | grep real
tail -6 gpurun_out/y_cplx.err
