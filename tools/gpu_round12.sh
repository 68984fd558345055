#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/microbench.py --what ivf > gpurun_out/micro_ivf5.jsonl 2> gpurun_out/micro_ivf5.err; cat gpurun_out/micro_ivf5.jsonl; tail -3 gpurun_out/micro_ivf5.err
