#!/bin/bash
# round 2, run V: topk = 1 instantiation of k_scan_stream32; bench hygiene (reference arm with real steps, prepared-subset leg)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/latency.py 2>&1 | tail -1 | tee -a gpurun_out/r02_latency_v.jsonl
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --linear-n 64000000 > gpurun_out/r02_bench_v.json 2> gpurun_out/r02_bench_v.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_v.json"))
print(d["value"], d["e2e"], d["roofline"]["frac"], d["roofline"]["traffic"], d["recall_at_1"])
print(json.dumps(d.get("roofline_linear_scan")))
print(json.dumps(d.get("subset_search")))
for x in d["sharded_large"]: print(json.dumps(x)[:700])
PY
tail -5 gpurun_out/r02_bench_v.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | cut -c1-700
