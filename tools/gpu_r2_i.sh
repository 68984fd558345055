#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_stream32 -s 5 -c 1 -f -o gpurun_out/r02_c5shape_stream python tools/phase_clocks.py --n 20000000 --nlist 10486 --persist 0 --ctas 1 --d 128 > gpurun_out/ncu_i1.log 2>&1; tail -2 gpurun_out/ncu_i1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_persist32 -s 2 -c 1 -f -o gpurun_out/r02_c2_persist python tools/phase_clocks.py --n 1000000 --nlist 1000 --d 128 --batch 8192 --split 0 --persist 1 > gpurun_out/ncu_i2.log 2>&1; tail -2 gpurun_out/ncu_i2.log
ls -la gpurun_out/*.ncu-rep | tail -3
