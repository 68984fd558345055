#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "opq_rotation or state_exchange or edge_cases or golden_query" 2>&1 | tail -4
timeout 300 python tools/build_large.py --config C4s 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --linear-n 0 --quick --no-large 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e'])"
