#!/bin/bash
# compute-sanitizer over the scan engines' tests (small sizes only).  Round 2: the persistent kernel (mbarrier hand-off), the
# fused multi-CTA query with the merge by the last CTA, the topk = 1 instantiations.  racecheck is ~100x slower than the
# other tools: it gets the three tests that cover the new synchronisation.
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 50 python -m pytest tests/test_gpu_parity.py -x -q -k "persistent_batch_kernel_exact or fused_coarse or golden_query_ivf or m64_ivf or two_phase or linear_auto_dispatch" > gpurun_out/r02_sanitize_$tool.log 2>&1
  tail -4 gpurun_out/r02_sanitize_$tool.log
done
timeout 900 compute-sanitizer --tool racecheck --print-limit 200 python -m pytest tests/test_gpu_parity.py -x -q -k "persistent_batch_kernel_exact or fused_coarse or m64_ivf" > gpurun_out/r02_sanitize_racecheck.log 2>&1
tail -4 gpurun_out/r02_sanitize_racecheck.log
