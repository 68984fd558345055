#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/r02_launches_bench_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --linear-n 0 --quick > gpurun_out/ncu_e0.log 2>&1; tail -1 gpurun_out/ncu_e0.log | cut -c1-200
bash tools/gpu_sanitize.sh
