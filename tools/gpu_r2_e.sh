#!/bin/bash
set -x
mkdir -p gpurun_out
for c in 0 1; do
RII_LARGE_CTAS=$c timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --linear-n 0 > gpurun_out/r02_bench_e$c.json 2> gpurun_out/r02_bench_e$c.err; python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_e$c.json"))
for x in d["sharded_large"]: print(json.dumps(x)[:1200])
print(json.dumps(d.get("subset_search")))
PY
tail -3 gpurun_out/r02_bench_e$c.err
done
