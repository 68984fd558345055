#!/bin/bash
# round 2, run P (session 2 re-entry): the state of the tree measured again -- bench both arms, launch list, ncu --set full of the
# persistent IVF kernel (C2) and of the C5-shaped sharded scan, phase clocks
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_BENCH_ours_p.json 2> gpurun_out/r02_BENCH_ours_p.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_BENCH_ours_p.json"))
    for k in ("value", "ms_per_step", "e2e", "roofline", "roofline_linear_scan", "subset_search", "cpu_baseline", "kernel_ms", "gpu_launches", "clocks"):
        print(k, json.dumps(d.get(k))[:700])
    for x in d.get("sharded_large") or []:
        print(json.dumps(x)[:1200])
except Exception as ex:
    print("bench failed", ex)
PY
tail -5 gpurun_out/r02_BENCH_ours_p.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_BENCH_reference_p.json 2> gpurun_out/r02_BENCH_reference_p.err; cut -c1-400 gpurun_out/r02_BENCH_reference_p.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_p.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --linear-n 0 --no-large > gpurun_out/ncu_p0.log 2>&1; tail -2 gpurun_out/ncu_p0.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_persist32 -s 2 -c 1 -f -o gpurun_out/r02_c2_persist python tools/phase_clocks.py --n 1000000 --nlist 1000 --d 128 --batch 8192 --split 0 --persist 1 > gpurun_out/ncu_p1.log 2>&1; tail -2 gpurun_out/ncu_p1.log | cut -c1-600
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_ -s 3 -c 1 -f -o gpurun_out/r02_c5shape_scan python tools/phase_clocks.py --n 20000000 --nlist 10486 > gpurun_out/ncu_p2.log 2>&1; tail -2 gpurun_out/ncu_p2.log | cut -c1-900
ls -la gpurun_out/*.ncu-rep | tail -3
