#!/bin/bash
# compute-sanitizer over the scan engines' tests (memcheck + racecheck + synccheck); small sizes only
set -x
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -k "(skew_kernel_matches_oracle and not 300001 and not 70001) or ivf_skew_kernel or fused_coarse or golden_query_ivf or m64_ivf or two_phase" > gpurun_out/sanitize_$tool.log 2>&1
  tail -12 gpurun_out/sanitize_$tool.log
done
