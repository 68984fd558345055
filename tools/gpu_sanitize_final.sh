#!/bin/bash
# memcheck + synccheck over the final tree's new paths (single calls with the query in the kernel parameters, topk = 1 merges,
# last-CTA merge, pipelined host batches, randomised shapes)
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  timeout 700 compute-sanitizer --tool $tool --print-limit 50 python -m pytest tests/test_gpu_parity.py -x -q -k "randomised or linear_auto_dispatch or golden_query or persistent_batch_kernel_exact or edge_cases" > gpurun_out/r02_sanitize_final_$tool.log 2>&1
  tail -4 gpurun_out/r02_sanitize_final_$tool.log
done
