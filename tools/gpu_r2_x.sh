#!/bin/bash
# round 2, run X: OR-merge of drain / first blocks in the persistent kernel; two-chunk host pipeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/phase_clocks.py --n 20000000 --nlist 10486 --persist 1 2>&1 | tail -1 | cut -c1-900 | tee -a gpurun_out/r02_phase_clocks_x.jsonl
timeout 300 python tools/phase_clocks.py --n 1000000 --nlist 1000 --d 128 --batch 8192 --split 0 --persist 1 2>&1 | tail -1 | cut -c1-900 | tee -a gpurun_out/r02_phase_clocks_x.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --linear-n 0 --quick --no-large 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', d['value'], d['e2e'], d['roofline']['frac'])"
