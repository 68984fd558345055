#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/microbench.py --what ivf > gpurun_out/micro_ivf7.jsonl 2> gpurun_out/micro_ivf7.err; cat gpurun_out/micro_ivf7.jsonl; tail -3 gpurun_out/micro_ivf7.err
timeout 300 python tools/microbench.py --n 64000000 --what linear --scan-kernel 2 --reps 5 > gpurun_out/micro_64M_sk2i.jsonl 2> gpurun_out/micro_sk2i.err; cat gpurun_out/micro_64M_sk2i.jsonl; tail -3 gpurun_out/micro_sk2i.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours21.json 2> gpurun_out/bench_ours21.err; cat gpurun_out/bench_ours21.json | cut -c1-900; tail -3 gpurun_out/bench_ours21.err
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -k "skew_kernel_matches_oracle or ivf_skew_kernel or fused_coarse" > gpurun_out/sanitize_racecheck2.log 2>&1; tail -8 gpurun_out/sanitize_racecheck2.log
