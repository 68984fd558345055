#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "persistent_batch" 2>&1 | tail -25
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --linear-n 0 --quick > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_g.json"))
print(d["value"], d["e2e"], d["roofline"]["frac"], d["recall_at_1"])
for x in d["sharded_large"]: print(json.dumps(x)[:1200])
PY
tail -5 gpurun_out/r02_bench_g.err
