#!/bin/bash
# 2-GPU validation (gpurun --gpus 2): bench.py under torchrun (replicated C2 + sharded C5 / C4 legs over NCCL), the multi-process
# shard parity worker (linear, IVF, IVF + target_ids against the unsharded oracle), and a 1-GPU bench line from the same tree
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_scale_n2.json 2> gpurun_out/r02_scale_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_scale_n2.json").read().strip().splitlines()[-1])
    print("N=2", d["value"], d["ms_per_step"], d["e2e"], d["recall_at_1"], d["config"]["parallelism"])
    print(d.get("replicas_with_all_gather"))
    for x in d["sharded_large"]: print(json.dumps(x)[:700])
except Exception as ex:
    print("failed", ex)
PY
tail -3 gpurun_out/r02_scale_n2.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/_shard_gpu_worker.py 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --linear-n 0 --quick --no-large 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', d['value'], d['e2e'], d['recall_at_1'])"
