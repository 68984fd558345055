#!/bin/bash
# last run of round 2: the GPU test-suite, smoke(), and both bench arms from the final tree
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_BENCH_ours_final.json 2> gpurun_out/r02_BENCH_ours_final.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_BENCH_ours_final.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "e2e", "roofline_linear_scan", "subset_search", "structured_data", "cpu_baseline"):
        print(k, json.dumps(d.get(k))[:500])
    print("roofline", d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline"]["hbm"]["frac"])
    for x in d.get("sharded_large") or []:
        print(x["workload"][:3], x["scan_kernel_ms"], x["scan_frac_of_hbm_peak"], x["queries_per_s"], x["parity_vs_oracle_at_full_scale"][:3])
except Exception as ex:
    print("bench failed", ex)
PY
tail -3 gpurun_out/r02_BENCH_ours_final.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_BENCH_reference_final.json 2> gpurun_out/r02_BENCH_reference_final.err; cut -c1-200 gpurun_out/r02_BENCH_reference_final.json
