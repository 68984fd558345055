#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "skew or fused or dual or stream" 2>&1 | tail -5
timeout 300 python tools/microbench.py --what ivf --scan-kernel 4 > gpurun_out/micro_ivf_sk4_c0.jsonl 2> gpurun_out/micro_ivf_sk4_c0.err; cat gpurun_out/micro_ivf_sk4_c0.jsonl; tail -3 gpurun_out/micro_ivf_sk4_c0.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sk4.json 2> gpurun_out/bench_sk4.err; cut -c1-700 gpurun_out/bench_sk4.json; tail -3 gpurun_out/bench_sk4.err
