#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/microbench.py --n 32000000 --m 64 --what linear --scan-kernel 4 --reps 5 > gpurun_out/micro_lin_m64.jsonl 2> gpurun_out/micro_lin_m64.err; cat gpurun_out/micro_lin_m64.jsonl; tail -3 gpurun_out/micro_lin_m64.err
timeout 600 python tools/configs_bench.py --only C4-scaled,C2,C5-scaled > gpurun_out/configs2.jsonl 2> gpurun_out/configs2.err; python -c "
import json
for l in open('gpurun_out/configs2.jsonl'):
    d=json.loads(l); print(d['config'], d['gpu']['queries_per_s'], d['gpu']['code_GBps_algorithmic'], d['gpu']['kernel_ms_per_batch'], d['reference'].get('ids_and_fp32_bits_identical_vs_strict_build'), d.get('speedup_vs_reference_per_query'))
"; tail -3 gpurun_out/configs2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sk4.json 2> gpurun_out/bench_sk4.err; cut -c1-400 gpurun_out/bench_sk4.json; tail -3 gpurun_out/bench_sk4.err
