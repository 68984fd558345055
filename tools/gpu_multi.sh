#!/bin/bash
# usage: gpurun --gpus G -- 'bash tools/gpu_multi.sh G'
set -x
G=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541 tests/_shard_gpu_worker.py 2>&1 | tail -15
for n in 1 $G; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  tail -c 1800 gpurun_out/scale_n$n.json; tail -5 gpurun_out/scale_n$n.err
done
