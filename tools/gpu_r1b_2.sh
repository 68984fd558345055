#!/bin/bash
# round 1, session 2: first run of the v4 (register streaming over skew64) engine -- parity, micro numbers, bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "skew or fused or dual or stream" 2>&1 | tail -15
for nw in 12 8; do
timeout 300 python tools/microbench.py --n 64000000 --what linear --scan-kernel 4 --stream-warps $nw --reps 5 > gpurun_out/micro_lin_sk4_$nw.jsonl 2> gpurun_out/micro_lin_sk4_$nw.err; cat gpurun_out/micro_lin_sk4_$nw.jsonl; tail -3 gpurun_out/micro_lin_sk4_$nw.err
timeout 300 python tools/microbench.py --what ivf --scan-kernel 4 --stream-warps $nw > gpurun_out/micro_ivf_sk4_$nw.jsonl 2> gpurun_out/micro_ivf_sk4_$nw.err; cat gpurun_out/micro_ivf_sk4_$nw.jsonl; tail -3 gpurun_out/micro_ivf_sk4_$nw.err
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sk4.json 2> gpurun_out/bench_sk4.err; cut -c1-1300 gpurun_out/bench_sk4.json; tail -3 gpurun_out/bench_sk4.err
