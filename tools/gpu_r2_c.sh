#!/bin/bash
# round 2, run C: new bench.py on one GPU (C2 headline, C3 subset leg, N=1B linear scan, C5/C4-shaped sharded legs at G=1)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "state_exchange or shards_on_one or cost" 2>&1 | tail -5
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench_c.json"))
    for k in ("value", "ms_per_step", "e2e", "roofline", "roofline_linear_scan", "subset_search", "sharded_large", "cpu_baseline", "kernel_ms"):
        print(k, json.dumps(d.get(k))[:900])
except Exception as ex:
    print("bench failed", ex)
PY
tail -5 gpurun_out/r02_bench_c.err
