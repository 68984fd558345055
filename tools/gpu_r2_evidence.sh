#!/bin/bash
# Round-2 evidence run (one B200): GPU test-suite, both bench arms, the ncu launch list of the bench command, ncu --set full
# captures of the hot kernels, phase clocks, single-call latency.  Everything lands in gpurun_out/; tools/ncu_summary.py and
# tools/ncu_source_hot.py turn the .ncu-rep files into the text files kept under profiles/.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu_final.log
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_BENCH_ours_final.json 2> gpurun_out/r02_BENCH_ours_final.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_BENCH_ours_final.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "e2e", "roofline", "roofline_linear_scan", "subset_search", "structured_data", "cpu_baseline", "gpu_launches", "clocks"):
        print(k, json.dumps(d.get(k))[:600])
    for x in d.get("sharded_large") or []:
        print(json.dumps(x)[:700])
except Exception as ex:
    print("bench failed", ex)
PY
tail -3 gpurun_out/r02_BENCH_ours_final.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_BENCH_reference_final.json 2> gpurun_out/r02_BENCH_reference_final.err; cut -c1-300 gpurun_out/r02_BENCH_reference_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --linear-n 0 --no-large --quick > gpurun_out/ncu_e0.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_persist32 -s 2 -c 1 -f -o gpurun_out/r02_c2_persist_final python tools/phase_clocks.py --n 1000000 --nlist 1000 --d 128 --batch 8192 --split 0 --persist 1 > gpurun_out/ncu_e1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_persist32 -s 2 -c 1 -f -o gpurun_out/r02_c5shape_persist_final python tools/phase_clocks.py --n 20000000 --nlist 10486 --persist 1 > gpurun_out/ncu_e2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_stream32 -s 5 -c 1 -f -o gpurun_out/r02_c4shape_stream_final python tools/phase_clocks.py --n 12500000 --nlist 10000 --d 128 --m 64 --persist 0 > gpurun_out/ncu_e3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_stream32 -s 3 -c 1 -f -o gpurun_out/r02_linear_stream_final python tools/microbench.py --n 32000000 --what linear --reps 3 > gpurun_out/ncu_e4.log 2>&1
ls -la gpurun_out/*final*.ncu-rep
rm -f gpurun_out/r02_phase_clocks_final.jsonl
timeout 300 python tools/phase_clocks.py --n 20000000 --nlist 10486 --persist 1 2>&1 | tail -1 | cut -c1-900 >> gpurun_out/r02_phase_clocks_final.jsonl
timeout 300 python tools/phase_clocks.py --n 1000000 --nlist 1000 --d 128 --batch 8192 --split 0 --persist 1 2>&1 | tail -1 | cut -c1-900 >> gpurun_out/r02_phase_clocks_final.jsonl
timeout 300 python tools/phase_clocks.py --n 12500000 --nlist 10000 --d 128 --m 64 --persist 0 2>&1 | tail -1 | cut -c1-900 >> gpurun_out/r02_phase_clocks_final.jsonl
cat gpurun_out/r02_phase_clocks_final.jsonl | cut -c1-700
timeout 300 python tools/latency.py 2>&1 | tail -1 > gpurun_out/r02_latency_final.json; cut -c1-900 gpurun_out/r02_latency_final.json
