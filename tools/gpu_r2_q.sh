#!/bin/bash
# round 2, run Q: persistent kernel v2 (mbarrier hand-off, run-based walk, topk = 1 instantiation, packed-fp32 table build)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "persistent or split_coarse or id_range_shards or fused_coarse or golden_query_ivf" 2>&1 | tail -8
timeout 300 python tools/phase_clocks.py --n 20000000 --nlist 10486 --persist 1 2>&1 | tail -1 | cut -c1-900 | tee -a gpurun_out/r02_phase_clocks.jsonl
timeout 300 python tools/phase_clocks.py --n 1000000 --nlist 1000 --d 128 --batch 8192 --split 0 --persist 1 2>&1 | tail -1 | cut -c1-900 | tee -a gpurun_out/r02_phase_clocks.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --linear-n 0 --quick > gpurun_out/r02_bench_q.json 2> gpurun_out/r02_bench_q.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_q.json"))
print(d["value"], d["e2e"], d["roofline"]["frac"], d["recall_at_1"])
for x in d["sharded_large"]: print(json.dumps(x)[:1200])
PY
tail -5 gpurun_out/r02_bench_q.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_persist32 -s 2 -c 1 -f -o gpurun_out/r02_c2_persist_v2 python tools/phase_clocks.py --n 1000000 --nlist 1000 --d 128 --batch 8192 --split 0 --persist 1 > gpurun_out/ncu_q1.log 2>&1; tail -1 gpurun_out/ncu_q1.log | cut -c1-200
