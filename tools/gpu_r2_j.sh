#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for ps in 0 1; do
timeout 300 python tools/phase_clocks.py --n 20000000 --nlist 10486 --persist $ps --ctas 1 2>&1 | tail -1 | cut -c1-900
timeout 300 python tools/phase_clocks.py --n 20000000 --nlist 10486 --persist $ps --ctas 0 2>&1 | tail -1 | cut -c1-900
timeout 300 python tools/phase_clocks.py --n 1000000 --nlist 1000 --d 128 --batch 8192 --split 0 --persist $ps 2>&1 | tail -1 | cut -c1-900
done
