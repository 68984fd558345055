#!/bin/bash
# usage: gpurun --gpus G -- 'bash tools/gpu_multi2.sh G'
set -x
G=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541 tests/_shard_gpu_worker.py > gpurun_out/shard_worker_g$G.log 2>&1; tail -3 gpurun_out/shard_worker_g$G.log; cat gpurun_out/shard_worker_rank*.log 2>/dev/null | tail -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $G --steps 20 --warmup 3 > gpurun_out/SCALE_n${G}_replica.json 2> gpurun_out/SCALE_n${G}_replica.err; tail -c 1500 gpurun_out/SCALE_n${G}_replica.json; tail -3 gpurun_out/SCALE_n${G}_replica.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $G --steps 20 --warmup 3 --shard > gpurun_out/SCALE_n${G}_shard.json 2> gpurun_out/SCALE_n${G}_shard.err; tail -c 1500 gpurun_out/SCALE_n${G}_shard.json; tail -3 gpurun_out/SCALE_n${G}_shard.err
