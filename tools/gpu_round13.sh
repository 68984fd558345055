#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/microbench.py --what ivf > gpurun_out/micro_ivf6.jsonl 2> gpurun_out/micro_ivf6.err; cat gpurun_out/micro_ivf6.jsonl; tail -3 gpurun_out/micro_ivf6.err
timeout 300 python tools/microbench.py --n 64000000 --what linear --scan-kernel 2 --reps 5 > gpurun_out/micro_64M_sk2h.jsonl 2> gpurun_out/micro_sk2h.err; cat gpurun_out/micro_64M_sk2h.jsonl; tail -3 gpurun_out/micro_sk2h.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours13.json 2> gpurun_out/bench_ours13.err; cat gpurun_out/bench_ours13.json | cut -c1-1700; tail -3 gpurun_out/bench_ours13.err
