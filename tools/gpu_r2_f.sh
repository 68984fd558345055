#!/bin/bash
set -x
mkdir -p gpurun_out
for c in 0 1; do
timeout 300 python tools/phase_clocks.py --n 20000000 --nlist 10486 --ctas $c --shards 1 2>&1 | tail -2
done
timeout 300 python tools/phase_clocks.py --n 20000000 --nlist 10486 --ctas 1 --d 128 2>&1 | tail -1
timeout 300 python tools/phase_clocks.py --n 1000000 --nlist 1000 --ctas 0 --d 128 --batch 2048 --split 0 2>&1 | tail -1
