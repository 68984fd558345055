#!/bin/bash
# round 1, session 2: first run of the v3 (dual-stream FFMA2) engine -- parity, micro numbers, bench
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "skew or fused or dual" 2>&1 | tail -15
timeout 300 python tools/microbench.py --n 64000000 --what linear --scan-kernel 3 --reps 5 > gpurun_out/micro_lin_sk3.jsonl 2> gpurun_out/micro_lin_sk3.err; cat gpurun_out/micro_lin_sk3.jsonl; tail -3 gpurun_out/micro_lin_sk3.err
timeout 300 python tools/microbench.py --what ivf --scan-kernel 3 > gpurun_out/micro_ivf_sk3.jsonl 2> gpurun_out/micro_ivf_sk3.err; cat gpurun_out/micro_ivf_sk3.jsonl; tail -3 gpurun_out/micro_ivf_sk3.err
timeout 300 python tools/microbench.py --what ivf --scan-kernel 2 > gpurun_out/micro_ivf_sk2.jsonl 2> gpurun_out/micro_ivf_sk2.err; cat gpurun_out/micro_ivf_sk2.jsonl; tail -3 gpurun_out/micro_ivf_sk2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sk3.json 2> gpurun_out/bench_sk3.err; cut -c1-1200 gpurun_out/bench_sk3.json; tail -3 gpurun_out/bench_sk3.err
