#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/microbench.py --what ivf > gpurun_out/micro_ivf2.jsonl 2> gpurun_out/micro_ivf2.err; cat gpurun_out/micro_ivf2.jsonl; tail -3 gpurun_out/micro_ivf2.err
timeout 300 python tools/microbench.py --n 64000000 --what linear --scan-kernel 2 > gpurun_out/micro_64M_sk2d.jsonl 2> gpurun_out/micro_sk2d.err; cat gpurun_out/micro_64M_sk2d.jsonl; tail -3 gpurun_out/micro_sk2d.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours8.json 2> gpurun_out/bench_ours8.err; cat gpurun_out/bench_ours8.json; tail -3 gpurun_out/bench_ours8.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_scan_skew32 -s 4 -c 1 -o gpurun_out/prof_skew_ivf4 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_skew_ivf4.log 2>&1
ls -la gpurun_out | tail -8
