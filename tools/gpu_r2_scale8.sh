#!/bin/bash
# 8-GPU pre-flight of what the driver runs at round end: bench.py under torchrun with 8 ranks (C2 replicated + the C5 / C4 legs at
# their real scale: 1e9 / 1e8 codes over 8 id-range shards)
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_scale_n8.json 2> gpurun_out/r02_scale_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_scale_n8.json").read().strip().splitlines()[-1])
    print("N=8", d["value"], d["ms_per_step"], d["e2e"], d["config"]["parallelism"], d["clocks"])
    for x in d["sharded_large"]: print(json.dumps(x)[:1000])
except Exception as ex:
    print("failed", ex)
PY
tail -5 gpurun_out/r02_scale_n8.err | cut -c1-300
