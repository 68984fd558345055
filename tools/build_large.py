"""REAL sharded index build at BASELINE's C4 / C5 scale (SURVEY 8 row g): codes drawn as randint, everything
else is the reference's algorithm on the GPUs -- the reference's sampling shuffle (src/rii.h:115-124), PQk-means fit on the
sample (src/pqkmeans.cpp:46-133), K6 assignment of every local code (src/rii.h:335-359), device-side posting lists,
exchange of the list lengths -- followed by sharded IVF batches with the coarse phase split over the ranks.

  torchrun --nproc-per-node 8 tools/build_large.py --config C5 [--iter 2]

Prints one JSON line per config on rank 0 (kept under profiles/)."""
import argparse
import ctypes as C
import datetime
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rii_b200 import _capi, main, sharded  # noqa: E402

CONFIGS = {"C5": dict(D=96, M=32, nlist=65536, N=1000000000, B=1024), "C4": dict(D=128, M=64, nlist=10000, N=100000000, B=1024),
           "C4s": dict(D=128, M=64, nlist=1000, N=8000000, B=1024)}


def p(t):
    return C.c_void_p(t.data_ptr())


def run():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C4")
    ap.add_argument("--iter", type=int, default=2)
    ap.add_argument("--n", type=int, default=0, help="override the total number of codes")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=600))
    lib = _capi.lib()
    cfg = dict(CONFIGS[a.config])
    if a.n:
        cfg["N"] = a.n
    D, M, nlist, N, B = cfg["D"], cfg["M"], cfg["nlist"], cfg["N"], cfg["B"]
    bounds = sharded.shard_bounds(N, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    Nl = hi - lo
    rng = np.random.default_rng(7)
    cw = rng.random((M, 256, D // M), dtype=np.float32)
    out = {"config": a.config, "N": N, "per_gpu": Nl, "world": world, "D": D, "M": M, "nlist": nlist, "iter": a.iter}
    e = main.RiiCpp(cw, False, device=local, l2_variant=16)
    lib.rii_profile_enable(e._h, 1)
    # ---- codes (randint, on the device) ----
    t0 = time.time()
    codes = torch.empty((Nl, M), dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    chunk = 1 << 24
    for s0 in range(0, Nl, chunk):
        c = min(chunk, Nl - s0)
        codes[s0:s0 + c] = torch.randint(0, 256, (c, M), dtype=torch.uint8, device=dev, generator=gen)
    _capi.check(lib.rii_reserve(e._h, Nl))
    _capi.check(lib.rii_add_codes_dev(e._h, p(codes), Nl, 0))
    _capi.check(lib.rii_set_shard(e._h, lo, N))
    torch.cuda.synchronize()
    out["t_codes_s"] = round(time.time() - t0, 2)
    # ---- the reference's sampling shuffle (host, rank 0) and the sample, assembled over the shards ----
    t0 = time.time()
    ns = min(N, 100 * nlist)
    if rank == 0:
        ids = torch.from_numpy(sharded.reference_sample_ids(N, nlist)).to(dev)
    else:
        ids = torch.empty((ns,), dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(ids, 0)
    out["t_sampling_shuffle_s"] = round(time.time() - t0, 2)
    t0 = time.time()
    mine = (ids >= lo) & (ids < hi)
    sample = torch.zeros((ns, M), dtype=torch.uint8, device=dev)
    sample[mine] = codes[ids[mine] - lo]
    if world > 1:
        dist.all_reduce(sample, op=dist.ReduceOp.SUM)  # disjoint rows: the sum is the concatenation
    sample_h = sample.cpu().numpy()
    del sample
    eng = sharded.CudaShardEngine(e)
    centers = eng.fit_coarse(sample_h, nlist, a.iter)   # replicated, deterministic (src/pqkmeans.cpp:46-133)
    torch.cuda.synchronize()
    out["t_fit_s"] = round(time.time() - t0, 2)
    # ---- K6 over the local codes + device-side posting lists ----
    t0 = time.time()
    lib.rii_profile_reset(e._h)
    eng.set_coarse_centers(centers)
    torch.cuda.synchronize()
    out["t_assign_and_lists_s"] = round(time.time() - t0, 2)
    ms, n = C.c_double(0), C.c_int64(0)
    lib.rii_profile_get(e._h, b"assign", C.byref(ms), C.byref(n))
    out["assign_kernel_s"] = round(ms.value / 1e3, 3)
    out["assign_T_lookups_per_s"] = round(Nl * nlist * M / (ms.value * 1e-3) / 1e12, 3) if ms.value > 0 else None
    lens = torch.from_numpy(eng.list_lengths()).to(dev)
    if world > 1:
        g = torch.empty((world, nlist), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(g.view(-1), lens)
    else:
        g = lens[None]
    glob = g.sum(0, dtype=torch.int32)
    pre = g[:rank].sum(0, dtype=torch.int32) if rank else torch.zeros_like(lens)
    eng.set_global_lengths(glob.cpu().numpy(), pre.cpu().numpy())
    out["list_len_min_mean_max"] = [int(glob.min()), float(glob.float().mean()), int(glob.max())]
    # ---- parity of the build on a sample (rank 0): assignments of 1000 local rows against the oracle ----
    if rank == 0:
        try:
            from oracle import oracle as O
            rows = torch.from_numpy(np.sort(np.random.default_rng(1).choice(Nl, 1000, replace=False))).to(dev)
            cr = codes[rows].cpu().numpy()
            oa = O.assign(O.sym_matrices(cw), cr, centers)
            off, lid = e.posting_lists_csr()
            got = np.searchsorted(off, np.array([np.nonzero(lid == r)[0][0] for r in rows.cpu().numpy()[:50]]), side="right") - 1
            ga = e.assign(cr, centers)
            out["build_parity"] = "ok" if (np.array_equal(ga, oa) and np.array_equal(got, oa[:50])) else "MISMATCH"
        except Exception as ex:
            out["build_parity"] = "error: " + repr(ex)
    del codes
    torch.cuda.empty_cache()
    # ---- sharded IVF batches, coarse phase split over the ranks ----
    L0 = int(round(N / nlist))
    L, k = min(32 * L0, N), 1
    Q = torch.from_numpy(np.random.default_rng(5).random((B, D), dtype=np.float32)).to(dev)
    for it in range(3):
        res = sharded.sharded_query_split(eng, Q, k, L, dist if world > 1 else None, world, rank)
    torch.cuda.synchronize()
    lib.rii_profile_reset(e._h)
    if world > 1:
        dist.barrier()
    evs = []
    for it in range(10):
        a0, b0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        res = sharded.sharded_query_split(eng, Q, k, L, dist if world > 1 else None, world, rank)
        b0.record()
        evs.append((a0, b0))
    torch.cuda.synchronize()
    msb = sum(x.elapsed_time(y) for x, y in evs) / len(evs)
    lib.rii_profile_get(e._h, b"scan_ivf", C.byref(ms), C.byref(n))
    out.update({"L": L, "batch": B, "ms_per_batch": round(msb, 4), "queries_per_s": round(B / (msb * 1e-3), 1),
                "scan_kernel_ms": round(ms.value / max(n.value, 1), 4),
                "scan_GBps_per_gpu": round(B * (L / world * M + 4 * M * 256) / (ms.value / max(n.value, 1) * 1e-3) / 1e9, 1)})
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    run()
