#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_stream32 -s 2 -c 1 -o gpurun_out/prof_v4_ivf2 python tools/microbench.py --what ivf --scan-kernel 4 > gpurun_out/ncu_v4_ivf2.log 2>&1; tail -3 gpurun_out/ncu_v4_ivf2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_stream32 -s 3 -c 1 -o gpurun_out/prof_v4_lin2 python tools/microbench.py --n 64000000 --what linear --reps 2 --scan-kernel 4 > gpurun_out/ncu_v4_lin2.log 2>&1; tail -3 gpurun_out/ncu_v4_lin2.log
