#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/phase_clocks.py --n 20000000 --nlist 10486 --persist 1 2>&1 | tail -1
timeout 300 python tools/phase_clocks.py --n 20000000 --nlist 10486 --persist 1 --d 128 2>&1 | tail -1
timeout 300 python tools/phase_clocks.py --n 1000000 --nlist 1000 --d 128 --batch 8192 --split 0 --persist 1 2>&1 | tail -1
