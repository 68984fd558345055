#!/bin/bash
# ncu full captures of the v4 engine (linear N=64M, fused IVF C2 batch)
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_stream32 -s 3 -c 1 -o gpurun_out/prof_v4_linear python tools/microbench.py --n 64000000 --what linear --reps 2 --scan-kernel 4 > gpurun_out/ncu_v4_lin.log 2>&1; tail -3 gpurun_out/ncu_v4_lin.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_stream32 -s 2 -c 1 -o gpurun_out/prof_v4_ivf python tools/microbench.py --what ivf --scan-kernel 4 > gpurun_out/ncu_v4_ivf.log 2>&1; tail -3 gpurun_out/ncu_v4_ivf.log
ls -la gpurun_out/*.ncu-rep
