#!/bin/bash
# round 2, run W (2 GPUs): pipelined host batches; bench.py under torchrun at N = 2 (replicated C2 + sharded C5 / C4 legs over NCCL)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --linear-n 0 --quick --no-large 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', d['value'], d['e2e'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_scale_n2_w.json 2> gpurun_out/r02_scale_n2_w.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_scale_n2_w.json").read().strip().splitlines()[-1])
    print("N=2", d["value"], d["ms_per_step"], d["e2e"], d["config"]["parallelism"])
    for x in d["sharded_large"]: print(json.dumps(x)[:900])
except Exception as ex:
    print("failed", ex)
PY
tail -5 gpurun_out/r02_scale_n2_w.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/_shard_gpu_worker.py 2>&1 | tail -3
