#!/bin/bash
# round 2, run A: parity after the restructure (v4-only engine, device-side list build, streaming assignment), K6 timing
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/microbench.py --what assign --n 1000000 > gpurun_out/r02_micro_assign.jsonl 2> gpurun_out/r02_micro_assign.err; cat gpurun_out/r02_micro_assign.jsonl; tail -3 gpurun_out/r02_micro_assign.err
timeout 300 python tools/microbench.py --what assign --n 1000000 --m 64 > gpurun_out/r02_micro_assign_m64.jsonl 2> gpurun_out/r02_micro_assign_m64.err; cat gpurun_out/r02_micro_assign_m64.jsonl; tail -3 gpurun_out/r02_micro_assign_m64.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; cut -c1-600 gpurun_out/r02_bench_a.json; tail -3 gpurun_out/r02_bench_a.err
