#!/bin/bash
# every BASELINE config (C4 / C5 scaled to one box) through the C ABI beside the unmodified reference: timing (fast build) and
# bit-exact parity (strict build)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 1500 python tools/configs_bench.py --nref 60 > gpurun_out/r02_configs_C1_C5.jsonl 2> gpurun_out/r02_configs.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_configs_C1_C5.jsonl"):
    d = json.loads(l)
    print(d["config"], d["gpu"]["queries_per_s"], d["gpu"]["single_query_call_us"], json.dumps(d.get("reference"))[:420], d.get("speedup_vs_reference_per_query"))
PY
tail -3 gpurun_out/r02_configs.err
