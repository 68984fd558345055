#!/bin/bash
# One gpurun call: parity tests, bench (both arms), ncu launch list + full capture, kernel microbench.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 3000 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
python tools/microbench.py --n 64000000 > gpurun_out/micro_64M.jsonl 2> gpurun_out/micro.err; cat gpurun_out/micro_64M.jsonl; tail -3 gpurun_out/micro.err
python tools/microbench.py --n 1000000 --what linear > gpurun_out/micro_1M.jsonl 2>> gpurun_out/micro.err; cat gpurun_out/micro_1M.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --batch 8192 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scan_ivf -s 3 -c 2 -o gpurun_out/prof_scan_ivf python bench.py --steps 3 --warmup 3 --no-cpu-baseline --batch 8192 > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scan_linear -s 6 -c 1 -o gpurun_out/prof_scan_linear python tools/microbench.py --n 64000000 --what linear --reps 2 > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out
