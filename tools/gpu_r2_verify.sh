#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --linear-n 0 --no-large 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value']); s=d['subset_search']; print(s['linear_queries_per_s'], s['ivf_queries_per_s'], s['ivf_prepared_subset_queries_per_s'])"
