#!/bin/bash
# round 2, run R: persistent kernel v2 without the prefetch logic in the issue path
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "persistent or split_coarse or id_range_shards or fused_coarse" 2>&1 | tail -3
timeout 300 python tools/phase_clocks.py --n 20000000 --nlist 10486 --persist 1 2>&1 | tail -1 | cut -c1-900 | tee -a gpurun_out/r02_phase_clocks_r.jsonl
timeout 300 python tools/phase_clocks.py --n 1000000 --nlist 1000 --d 128 --batch 8192 --split 0 --persist 1 2>&1 | tail -1 | cut -c1-900 | tee -a gpurun_out/r02_phase_clocks_r.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --linear-n 0 --quick > gpurun_out/r02_bench_r.json 2> gpurun_out/r02_bench_r.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_r.json"))
print(d["value"], d["e2e"], d["roofline"]["frac"], d["recall_at_1"])
for x in d["sharded_large"]: print(json.dumps(x)[:700])
PY
tail -5 gpurun_out/r02_bench_r.err
