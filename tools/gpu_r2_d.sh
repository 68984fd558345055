#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "split_coarse" 2>&1 | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --linear-n 0 --quick > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_d.json"))
for x in d["sharded_large"]: print(json.dumps(x)[:3000])
PY
tail -5 gpurun_out/r02_bench_d.err
