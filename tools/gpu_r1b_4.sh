#!/bin/bash
# v4 engine with two CTAs per SM for per-query IVF batches: parity, micro numbers, bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "skew or fused or dual or stream" 2>&1 | tail -15
for c in 0 1; do
timeout 300 python tools/microbench.py --what ivf --scan-kernel 4 --stream-ctas $c > gpurun_out/micro_ivf_sk4_c$c.jsonl 2> gpurun_out/micro_ivf_sk4_c$c.err; cat gpurun_out/micro_ivf_sk4_c$c.jsonl; tail -3 gpurun_out/micro_ivf_sk4_c$c.err
done
timeout 300 python tools/microbench.py --n 64000000 --what linear --scan-kernel 4 --reps 5 > gpurun_out/micro_lin_sk4.jsonl 2> gpurun_out/micro_lin_sk4.err; cat gpurun_out/micro_lin_sk4.jsonl; tail -3 gpurun_out/micro_lin_sk4.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sk4.json 2> gpurun_out/bench_sk4.err; cut -c1-1300 gpurun_out/bench_sk4.json; tail -3 gpurun_out/bench_sk4.err
