"""One measurement line per BASELINE.json config (C1..C5; the big ones scaled to what one GPU box and the CPU reference
can build in seconds -- stated in each line), through the public API (rii_b200.main.RiiCpp, i.e. the C ABI), with the
UNMODIFIED reference (oracle/_ref/fast_*: its own sources and flags) built on the same codes, timed beside it and used as
the parity check (ids equal; distances within 1e-5 relative, the BASELINE contract against the -Ofast build).
Not a bench line: feeds profiles/ and DESIGN.md.      python tools/configs_bench.py [--only C1,C3]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rii_b200 import _capi, main  # noqa: E402

CONFIGS = [
    # name, N, D, M, nlist, method, topk, L, subset S, GPU batch, note
    dict(name="C1", N=10000, D=128, M=32, nlist=0, method="linear", topk=3, L=0, S=0, B=1024,
         note="README example: linear PQ scan, topk=3"),
    dict(name="C2", N=1000000, D=128, M=32, nlist=1000, method="ivf", topk=1, L=32000, S=0, B=8192,
         note="the bench.py workload (here: random codes)"),
    dict(name="C3-linear", N=1000000, D=128, M=32, nlist=1000, method="linear", topk=10, L=0, S=100000, B=256,
         note="subset search, target_ids = 100k random, linear"),
    dict(name="C3-ivf", N=1000000, D=128, M=32, nlist=1000, method="ivf", topk=10, L=32000, S=100000, B=256,
         note="subset search, target_ids = 100k random, ivf"),
    dict(name="C4-scaled", N=1000000, D=128, M=64, nlist=100, method="ivf", topk=1, L=320000, S=0, B=1024,
         note="C4 is N=100M, nlist=10^4 on 8 GPUs; scaled to N=1M, nlist=100 (same list length 10^4, L = 32 L0)"),
    dict(name="C5-scaled", N=2000000, D=96, M=32, nlist=2048, method="ivf", topk=10, L=31250, S=0, B=1024,
         note="C5 is N=1B, nlist=65536, batch 1024 on 8 GPUs; scaled to N=2M, nlist=2048 (> 1024: coarse ranking in the "
              "warps' top-k lists), L = 32 L0, batch 1024, D=96 (Ds=3)"),
]


def run(cfg, nref):
    lib = _capi.lib()
    dev = torch.device("cuda", 0)
    N, D, M, Ks = cfg["N"], cfg["D"], cfg["M"], 256
    rng = np.random.default_rng(1000 + [c["name"] for c in CONFIGS].index(cfg["name"]))
    cw = rng.random((M, Ks, D // M), dtype=np.float32)
    codes = rng.integers(0, 256, (N, M), dtype=np.uint8)
    Q = rng.random((cfg["B"], D), dtype=np.float32)
    tids = np.sort(rng.choice(N, cfg["S"], replace=False)).astype(np.int64) if cfg["S"] else np.empty(0, np.int64)
    e = main.RiiCpp(cw, False, l2_variant=16)
    e.add_codes(codes, False)
    t0 = time.time()
    if cfg["nlist"]:
        e.reconfigure(cfg["nlist"], 5)
    torch.cuda.synchronize()
    t_build = time.time() - t0
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    sp = C.c_void_p(st.cuda_stream)
    B, k = cfg["B"], cfg["topk"]
    dQ = torch.from_numpy(Q).to(dev)
    dT = torch.from_numpy(tids).to(dev) if cfg["S"] else None
    oi = torch.empty((B, k), dtype=torch.int64, device=dev)
    od = torch.empty((B, k), dtype=torch.float32, device=dev)
    oc = torch.empty((B,), dtype=torch.int32, device=dev)
    meth = 1 if cfg["method"] == "ivf" else 0

    def step():
        _capi.check(lib.rii_query_batch_dev(e._h, C.c_void_p(dQ.data_ptr()), B, k,
                                            C.c_void_p(dT.data_ptr()) if dT is not None else None, cfg["S"], cfg["L"], meth,
                                            C.c_void_p(oi.data_ptr()), C.c_void_p(od.data_ptr()), C.c_void_p(oc.data_ptr()), sp))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    lib.rii_profile_enable(e._h, 1)
    lib.rii_profile_reset(e._h)
    evs = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        step()
        b.record(st)
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = min(a.elapsed_time(b) for a, b in evs)
    kern = {}
    for name in ("dtable", "scan_linear", "coarse_rank", "count_members", "plan", "scan_ivf", "merge"):
        m_, n_ = C.c_double(0), C.c_int64(0)
        lib.rii_profile_get(e._h, name.encode(), C.byref(m_), C.byref(n_))
        if n_.value:
            kern[name] = round(m_.value / 5, 4)
    lib.rii_profile_enable(e._h, 0)
    ids_gpu, d_gpu, cnt = oi.cpu().numpy(), od.cpu().numpy(), oc.cpu().numpy()
    # single-query latency through the reference's own call shape
    for q in Q[:10]:  # (the first calls allocate the pinned staging buffer and load the single-call kernel instantiations)
        e.query_ivf(q, k, tids, cfg["L"]) if meth else e.query_linear(q, k, tids)
    t1 = time.perf_counter()
    for q in Q[:50]:
        e.query_ivf(q, k, tids, cfg["L"]) if meth else e.query_linear(q, k, tids)
    lat = (time.perf_counter() - t1) / 50
    cand = cfg["L"] if meth else (cfg["S"] or N)
    line = {"config": cfg["name"], "note": cfg["note"], "N": N, "D": D, "M": M, "nlist": cfg["nlist"], "method": cfg["method"],
            "topk": k, "L": cfg["L"], "target_ids": cfg["S"], "data": "random codes / codewords / queries (throughput + parity)",
            "gpu": {"batch": B, "ms_per_batch": round(ms, 4), "queries_per_s": round(B / (ms * 1e-3), 1),
                    "code_GBps_algorithmic": round(B * cand * M / (ms * 1e-3) / 1e9, 1), "kernel_ms_per_batch": kern,
                    "single_query_call_us": round(lat * 1e6, 1), "index_build_s": round(t_build, 2)}}
    # ---- the reference on the same codes: its own build + reconfigure.  `fast` (-Ofast, what users run) is timed;
    # `strict` (the same sources without fast-math) is the parity check: under -Ofast the reference's own PQk-means can
    # drift (M = 64: the two builds of the reference disagree on the coarse centers), so ids are compared with strict.
    try:
        from oracle import ref as R
        if not R.available("fast") or not R.available("strict"):
            raise RuntimeError("oracle/_ref/{fast,strict}_* not built")
        r = R.Ref("fast")
        r.create(cw)
        r.add_codes(codes, False)
        tb = r.reconfigure(cfg["nlist"], 5) if cfg["nlist"] else 0.0
        out = r.time_queries(Q[:nref], k, cfg["method"], L=cfg["L"], tids=tids if cfg["S"] else None, warmup=2, return_ids=True)
        same_fast = sum(int(list(ids) == ids_gpu[i][:cnt[i]].tolist()) for i, ids in enumerate(out["ids"]))
        r.close()
        r = R.Ref("strict")
        r.create(cw)
        r.add_codes(codes, False)
        if cfg["nlist"]:
            r.reconfigure(cfg["nlist"], 5)
        same = bitexact = 0
        npar = min(nref, 40)
        for i in range(npar):
            q = Q[i]
            ids, dd = (r.query_ivf(q, k, tids, cfg["L"]) if meth else r.query_linear(q, k, tids if cfg["S"] else None))
            ok = list(ids) == ids_gpu[i][:cnt[i]].tolist()
            same += int(ok)
            bitexact += int(ok and np.array_equal(np.asarray(dd, np.float32).view(np.uint32), d_gpu[i][:cnt[i]].view(np.uint32)))
        r.close()
        line["reference"] = {"timed_build": "fast", "reconfigure_s": round(tb, 2),
                             "ms_per_query": round(out["seconds"] / out["n"] * 1e3, 4), "queries_timed": out["n"],
                             "threads": "OpenMP over %d cores (linear); QueryIvf is single-threaded" % (os.cpu_count() or 1),
                             "id_lists_identical_vs_fast_build": "%d/%d" % (same_fast, out["n"]),
                             "id_lists_identical_vs_strict_build": "%d/%d" % (same, npar),
                             "ids_and_fp32_bits_identical_vs_strict_build": "%d/%d" % (bitexact, npar)}
        line["speedup_vs_reference_per_query"] = round((out["seconds"] / out["n"]) / (ms * 1e-3 / B), 1)
    except Exception as ex:  # keep the GPU numbers
        line["reference"] = {"error": repr(ex)}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--nref", type=int, default=100)
    a = ap.parse_args()
    for cfg in CONFIGS:
        if a.only and cfg["name"] not in a.only.split(","):
            continue
        run(cfg, a.nref)
