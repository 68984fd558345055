#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/_shard_gpu_worker.py > gpurun_out/shard_worker.log 2>&1; tail -5 gpurun_out/shard_worker.log; cat gpurun_out/shard_worker_rank*.log 2>/dev/null | tail -30
timeout 300 python tools/microbench.py --what ivf > gpurun_out/micro_ivf.jsonl 2> gpurun_out/micro_ivf.err; cat gpurun_out/micro_ivf.jsonl; tail -3 gpurun_out/micro_ivf.err
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
