#!/bin/bash
mkdir -p gpurun_out
for b in 1 8; do
timeout 300 python tools/phase_clocks.py --n 1000000 --nlist 1000 --d 128 --batch $b --split 0 --persist 0 2>&1 | tail -1 | cut -c1-900
done
