#!/bin/bash
# single-GPU measurement set of round 1 (session 2, v4 engine): tests, both bench arms, microbench, ncu launch list + full captures
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/BENCH_ours.json 2> gpurun_out/BENCH_ours.err; cut -c1-400 gpurun_out/BENCH_ours.json; tail -3 gpurun_out/BENCH_ours.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/BENCH_ref.json 2> gpurun_out/BENCH_ref.err; cut -c1-300 gpurun_out/BENCH_ref.json
timeout 300 python tools/microbench.py --what ivf,assign --n 1000000 > gpurun_out/micro_ivf_final.jsonl 2>/dev/null; cat gpurun_out/micro_ivf_final.jsonl | cut -c1-500
timeout 300 python tools/microbench.py --n 64000000 --what linear > gpurun_out/micro_linear_final.jsonl 2>/dev/null; cat gpurun_out/micro_linear_final.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 8 -c 120 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --linear-n 0 > gpurun_out/ncu_l.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_scan_stream32 -s 4 -c 1 -o gpurun_out/prof_final_ivf python bench.py --steps 3 --warmup 3 --no-cpu-baseline --linear-n 0 > gpurun_out/ncu_f1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_scan_stream32 -s 3 -c 1 -o gpurun_out/prof_final_linear python tools/microbench.py --n 64000000 --what linear --reps 2 > gpurun_out/ncu_f2.log 2>&1
ls -la gpurun_out | tail -14
