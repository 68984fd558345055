#!/bin/bash
# round 2, run T: one launch per small call (fused multi-CTA query + merge by the last CTA); latency breakdown
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for n in 1000000 100000; do
timeout 300 python tools/latency.py --n $n --nlist $((n/1000)) 2>&1 | tail -1 | tee -a gpurun_out/r02_latency_t2.jsonl
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --linear-n 0 > gpurun_out/r02_bench_t.json 2> gpurun_out/r02_bench_t.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_t.json"))
print(d["value"], d["e2e"], d["roofline"]["frac"], d["recall_at_1"])
print(json.dumps(d.get("subset_search")))
for x in d["sharded_large"]: print(json.dumps(x)[:700])
PY
tail -5 gpurun_out/r02_bench_t.err
