#!/bin/bash
# round 2, run S: cooperative first table; single-query latency breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "persistent or split_coarse or id_range_shards or fused_coarse" 2>&1 | tail -3
timeout 300 python tools/phase_clocks.py --n 20000000 --nlist 10486 --persist 1 2>&1 | tail -1 | cut -c1-900 | tee -a gpurun_out/r02_phase_clocks_s.jsonl
timeout 300 python tools/phase_clocks.py --n 1000000 --nlist 1000 --d 128 --batch 8192 --split 0 --persist 1 2>&1 | tail -1 | cut -c1-900 | tee -a gpurun_out/r02_phase_clocks_s.jsonl
timeout 300 python tools/latency.py 2>&1 | tail -1 | tee gpurun_out/r02_latency_s.json
