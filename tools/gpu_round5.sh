#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours5.json 2> gpurun_out/bench_ours5.err; cat gpurun_out/bench_ours5.json; tail -3 gpurun_out/bench_ours5.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_scan_skew32 -s 4 -c 1 -o gpurun_out/prof_skew_ivf3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_skew_ivf3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 20 -c 200 --csv --log-file gpurun_out/launches_bench5.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench5.log 2>&1
ls -la gpurun_out
