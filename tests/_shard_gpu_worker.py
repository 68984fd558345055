"""Multi-GPU parity worker (run under torchrun on a box with >= 2 GPUs; tools/gpu_multi2.sh):
id-range sharded index over NCCL == unsharded oracle, for linear and IVF batches."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import O, synth  # noqa: E402
from rii_b200 import main, sharded  # noqa: E402


def run():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    D, M, Ks, N, nlist = 128, 32, 256, 200000, 100
    cw, codes, Q = synth(D, M, Ks, N, 8, seed=1234)
    e = main.RiiCpp(cw, False, device=local, l2_variant=16)
    centers = sharded.build_shard(e, codes, nlist, 2, rank, world)
    eng = sharded.CudaShardEngine(e)
    oc, oa = O.reconfigure(cw, codes, nlist, 2)
    assert np.array_equal(centers, oc), "sharded coarse centers differ from the oracle"
    offsets, ids = O.assign_to_lists(oa, nlist)
    Qd = torch.from_numpy(Q).to(dev)
    for method, topk, L in [("linear", 1, 0), ("linear", 50, 0), ("ivf", 1, 2000), ("ivf", 10, 6400), ("ivf", 100, 150),
                            ("ivf", 5, N)]:
        gi, gd, gc = sharded.sharded_query(eng, Qd, topk, L, method, dist, world)
        torch.cuda.synchronize()
        gi, gd, gc = gi.cpu().numpy(), gd.cpu().numpy(), gc.cpu().numpy()
        for b, q in enumerate(Q):
            T = O.dtable(q, cw, 16)
            exp = O.query_linear(T, codes, topk) if method == "linear" else O.query_ivf(T, codes, oc, offsets, ids, topk, L)
            n = int(gc[b])
            assert n == len(exp[0]), (method, topk, L, n, len(exp[0]))
            assert np.array_equal(gi[b, :n], exp[0]), (method, topk, L, gi[b, :n][:5], exp[0][:5])
            assert np.array_equal(gd[b, :n].view(np.uint32), exp[1].view(np.uint32)), (method, topk, L)
    # IVF + target_ids across shards (two-phase C ABI): uniform subsets and one concentrated in far lists (flagged re-run)
    rng = np.random.default_rng(5)
    far = np.argsort(O.adist_all(O.dtable(Q[0], cw, 16), oc))[-5:]
    far_ids = np.sort(np.concatenate([ids[offsets[no]:offsets[no + 1]] for no in far])).astype(np.int64)
    for tids, topk, L in [(np.sort(rng.choice(N, 20000, replace=False)).astype(np.int64), 10, 3200),
                          (np.sort(rng.choice(N, 3000, replace=False)).astype(np.int64), 1, 3000),
                          (far_ids, 3, 40)]:
        gi, gd, gc = sharded.sharded_query_subset(eng, Qd, topk, L, torch.from_numpy(tids).to(dev), dist, world, rank)
        torch.cuda.synchronize()
        gi, gd, gc = gi.cpu().numpy(), gd.cpu().numpy(), gc.cpu().numpy()
        for b, q in enumerate(Q):
            exp = O.query_ivf(O.dtable(q, cw, 16), codes, oc, offsets, ids, topk, L, tids)
            n = int(gc[b])
            assert n == len(exp[0]), ("subset", topk, L, n, len(exp[0]))
            assert np.array_equal(gi[b, :n], exp[0]), ("subset", topk, L, gi[b, :n][:5], exp[0][:5])
            assert np.array_equal(gd[b, :n].view(np.uint32), exp[1].view(np.uint32)), ("subset", topk, L)
    dist.barrier()
    if rank == 0:
        print("sharded GPU parity ok (linear, ivf, ivf + target_ids): world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        run()
    except BaseException:
        import traceback
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "shard_worker_rank%s.log" % os.environ.get("RANK", "x")), "w") as f:
            traceback.print_exc(file=f)
        raise
