#!/bin/bash
# 2-GPU validation: bench.py under torchrun (replicas + sharded legs over NCCL), the multi-process shard parity worker
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_scale_n2.json 2> gpurun_out/r02_scale_n2.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_scale_n2.json"))
    print(d["value"], d["ms_per_step"], d["e2e"], d["config"]["parallelism"])
    for x in d["sharded_large"]: print(json.dumps(x)[:1100])
except Exception as ex:
    print("failed", ex)
PY
tail -5 gpurun_out/r02_scale_n2.err
echo skip
