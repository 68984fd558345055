#!/bin/bash
# round 2, run B: parity after the query-side rewrite (sub-index subset search, coarse split, general path)
set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
