#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for sk in 1 2; do
  timeout 300 python tools/microbench.py --n 64000000 --what linear --scan-kernel $sk > gpurun_out/micro_64M_sk$sk.jsonl 2> gpurun_out/micro_sk$sk.err; cat gpurun_out/micro_64M_sk$sk.jsonl; tail -3 gpurun_out/micro_sk$sk.err
done
timeout 300 python tools/microbench.py --n 512000000 --what linear --scan-kernel 2 --reps 5 > gpurun_out/micro_512M_sk2.jsonl 2>> gpurun_out/micro_sk2.err; cat gpurun_out/micro_512M_sk2.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:skew -s 3 -c 1 -o gpurun_out/prof_scan_skew python tools/microbench.py --n 64000000 --what linear --reps 2 --scan-kernel 2 > gpurun_out/ncu_skew.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours2.json 2> gpurun_out/bench_ours2.err; cat gpurun_out/bench_ours2.json; tail -3 gpurun_out/bench_ours2.err
ls -la gpurun_out
