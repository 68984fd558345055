#!/bin/bash
# round 2, run Y: run-based issue logic in k_scan_stream32 (linear scans, C4, single calls); clustered-data leg
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/latency.py 2>&1 | tail -1 | cut -c1-1500 | tee -a gpurun_out/r02_latency_y.jsonl
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --linear-n 64000000 > gpurun_out/r02_bench_y.json 2> gpurun_out/r02_bench_y.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_y.json"))
print(d["value"], d["e2e"], d["roofline"]["frac"], d["recall_at_1"])
print(json.dumps(d.get("roofline_linear_scan")))
print(json.dumps(d.get("subset_search")))
print(json.dumps(d.get("clustered_data")))
for x in d["sharded_large"]: print(json.dumps(x)[:600])
PY
tail -5 gpurun_out/r02_bench_y.err
