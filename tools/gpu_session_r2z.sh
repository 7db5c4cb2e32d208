#!/bin/bash
# Round-2 session Z (1 GPU): ncu evidence of the native complex kernels on the complex twin of C2 -- one
# `--set full` capture of kry_spmv_csr_z's staged kernel and of zorth_kernel at k = 19 of the second cycle
# (launches alternate SpMV / Gram-Schmidt; 61 + 38 are skipped), read back as text here
set -u
mkdir -p gpurun_out
( time timeout 120 ncu --set full --clock-control none --import-source on -k regex:'zorth_kernel|zspmv_staged' -s 99 -c 2 -f \
    -o gpurun_out/z_cplx python tools/profile_cplx.py > gpurun_out/z_ncu.log 2>&1 ) 2>&1 | grep real
tail -3 gpurun_out/z_ncu.log
ncu -i gpurun_out/z_cplx.ncu-rep --page details > gpurun_out/z_cplx_details.txt 2>/dev/null
ncu -i gpurun_out/z_cplx.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active > gpurun_out/z_cplx_raw.csv 2>/dev/null
grep -n "zorth_kernel\|zspmv_staged\|  Duration\|DRAM Throughput\|Memory Throughput\|Registers Per\|Achieved Occupancy\|Theoretical Occupancy" gpurun_out/z_cplx_details.txt | head -30
cat gpurun_out/z_cplx_raw.csv | tail -4 | cut -c1-600
