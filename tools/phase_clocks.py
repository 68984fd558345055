"""Per-CTA phase clocks of the fused / split IVF scan kernel for an arbitrary shape (random codes, random lists).
python tools/phase_clocks.py --n 20000000 --d 96 --m 32 --nlist 10486 --batch 1024 --split 1 [--ctas 1]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rii_b200 import _capi, main  # noqa: E402


def p(t):
    return C.c_void_p(t.data_ptr())


def run():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20000000)
    ap.add_argument("--d", type=int, default=96)
    ap.add_argument("--m", type=int, default=32)
    ap.add_argument("--nlist", type=int, default=10486)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--lmul", type=int, default=32)
    ap.add_argument("--topk", type=int, default=1)
    ap.add_argument("--split", type=int, default=1)
    ap.add_argument("--ctas", type=int, default=0)
    ap.add_argument("--persist", type=int, default=1)
    ap.add_argument("--shards", type=int, default=1, help="pretend this GPU holds 1/shards of every list (global lengths = shards x local)")
    a = ap.parse_args()
    lib = _capi.lib()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    cw = rng.random((a.m, 256, a.d // a.m), dtype=np.float32)
    centers = rng.integers(0, 256, (a.nlist, a.m), dtype=np.uint8)
    e = main.RiiCpp(cw, False, l2_variant=16)
    codes = torch.randint(0, 256, (a.n, a.m), dtype=torch.uint8, device=dev)
    assign = torch.randint(0, a.nlist, (a.n,), dtype=torch.int32, device=dev)
    _capi.check(lib.rii_add_codes_dev(e._h, p(codes), a.n, 0))
    G = a.shards
    _capi.check(lib.rii_set_shard(e._h, 0, a.n * G))
    _capi.check(lib.rii_set_lists_dev(e._h, centers.ctypes.data_as(C.POINTER(C.c_uint8)), a.nlist, p(assign)))
    lens = torch.bincount(assign, minlength=a.nlist).to(torch.int32).cpu().numpy()
    glob = (lens * G).astype(np.int32)
    pre = np.zeros_like(glob)
    _capi.check(lib.rii_set_global_lengths(e._h, glob.ctypes.data_as(C.POINTER(C.c_int32)), pre.ctypes.data_as(C.POINTER(C.c_int32))))
    L0 = int(round(a.n * G / a.nlist))
    L = a.lmul * L0
    B, k = a.batch, a.topk
    Q = torch.rand((B, a.d), device=dev)
    e.set_option("stream_ctas", a.ctas)
    e.set_option("persist", a.persist)
    e.set_option("debug_clocks", 1)
    lib.rii_profile_enable(e._h, 1)
    w = _capi.check(lib.rii_coarse_width(e._h, L))
    ranked = torch.empty((B, w), dtype=torch.int32, device=dev)
    oi = torch.empty((B, k), dtype=torch.int64, device=dev)
    od = torch.empty((B, k), dtype=torch.float32, device=dev)
    oc = torch.empty((B,), dtype=torch.int32, device=dev)
    fl = torch.empty((B,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    sp = C.c_void_p(st.cuda_stream)
    for it in range(6):
        if it == 2:
            torch.cuda.synchronize()
            lib.rii_profile_reset(e._h)
        if a.split:
            _capi.check(lib.rii_coarse_rank_dev(e._h, p(Q), B, k, L, p(ranked), sp))
            _capi.check(lib.rii_query_ranked_dev(e._h, p(Q), B, k, L, p(ranked), p(oi), p(od), p(oc), p(fl), sp))
        else:
            _capi.check(lib.rii_query_batch_dev(e._h, p(Q), B, k, None, 0, L, 1, p(oi), p(od), p(oc), sp))
    torch.cuda.synchronize()
    clk = np.zeros((max(B, 296), 8), np.int64)
    _capi.check(lib.rii_debug_clocks(e._h, max(B, 296), clk.ctypes.data_as(C.POINTER(C.c_int64))))
    out = {"shape": vars(a), "L": L, "w": w}
    for name in ("coarse_rank", "scan_ivf"):
        m_, n_ = C.c_double(0), C.c_int64(0)
        lib.rii_profile_get(e._h, name.encode(), C.byref(m_), C.byref(n_))
        out[name + "_ms"] = round(m_.value / max(n_.value, 1), 4)
    if a.persist and B >= 296:  # persistent kernel: 16 counters per CTA (2 per 8-word row)
        c = clk.reshape(-1, 16)[:148]
        names = ["cons0_wait_full", "cons0_scan", "cons10_wait_full", "cons10_scan", "prod_wait_done", "prod_merge", "prod_table", "prod_coarse",
                 "prod_select", "prod_plan", "queries_per_cta"]
        nq = np.maximum(c[:, 10], 1)
        out["persist_cycles_per_query"] = {n: float((c[:, i] / nq).mean()) for i, n in enumerate(names[:10])}
        out["scan_GBps"] = round(B * (L / G * a.m + 4 * a.m * 256) / (out["scan_ivf_ms"] * 1e-3) / 1e9, 1)
        print(json.dumps(out))
        return
    t0 = clk[:, 0]
    out["cta_cycles_mean"] = {"table_built": float((clk[:, 4] - t0).mean()), "ready_to_scan": float((clk[:, 1] - t0).mean()),
                              "scan": float((clk[:, 2] - clk[:, 1]).mean()), "tail": float((clk[:, 3] - clk[:, 2]).mean()),
                              "total": float((clk[:, 3] - t0).mean())}
    alg = B * (L / G * a.m + 4 * a.m * 256)
    out["scan_GBps"] = round(alg / (out["scan_ivf_ms"] * 1e-3) / 1e9, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    run()
