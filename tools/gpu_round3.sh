#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/microbench.py --n 64000000 --what linear --scan-kernel 2 > gpurun_out/micro_64M_sk2b.jsonl 2> gpurun_out/micro_sk2b.err; cat gpurun_out/micro_64M_sk2b.jsonl; tail -3 gpurun_out/micro_sk2b.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours3.json 2> gpurun_out/bench_ours3.err; cat gpurun_out/bench_ours3.json; tail -3 gpurun_out/bench_ours3.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_scan_skew32 -s 3 -c 1 -o gpurun_out/prof_skew_linear python tools/microbench.py --n 64000000 --what linear --reps 2 --scan-kernel 2 > gpurun_out/ncu_skew_lin.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_scan_skew32 -s 4 -c 1 -o gpurun_out/prof_skew_ivf python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_skew_ivf.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out
