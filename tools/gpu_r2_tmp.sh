#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-large > gpurun_out/r02_scale_n2_c.json 2> gpurun_out/r02_scale_n2_c.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_scale_n2_c.json").read().strip().splitlines()[-1])
    print("N=2", d["value"], d["ms_per_step"], d["e2e"], d["recall_at_1"], d["roofline"]["launch_ms"], d["roofline"]["queries_per_launch"], d["config"]["parallelism"])
    print(d.get("replicas_with_all_gather"), d.get("replicas_without_all_gather"))
except Exception as ex:
    print("failed", ex)
PY
tail -3 gpurun_out/r02_scale_n2_c.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 --no-large --gather 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('gather', d['value'], d['e2e'], d['recall_at_1'], d.get('replicas_without_all_gather'))"
