#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "m64 or padded or split_coarse or assign_stream" 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_k.json 2> gpurun_out/r02_bench_k.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_k.json"))
print(d["value"], d["e2e"], d["roofline"]["frac"], d["roofline_linear_scan"].get("frac"))
print(json.dumps(d.get("subset_search")))
for x in d["sharded_large"]: print(json.dumps(x)[:1000])
PY
tail -3 gpurun_out/r02_bench_k.err
