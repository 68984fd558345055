"""Where a single-query call (the reference's call shape, src/main.cpp:17-27) spends its time on the C2 index:
Python wrapper (main.RiiCpp.query_ivf), the bare C-ABI call through ctypes with preallocated buffers, and the device time
of the kernels of one call (CUDA events inside the library).  python tools/latency.py [--n 1000000]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rii_b200 import _capi, main  # noqa: E402


def run():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000000)
    ap.add_argument("--nlist", type=int, default=1000)
    ap.add_argument("--topk", type=int, default=1)
    ap.add_argument("--calls", type=int, default=2000)
    a = ap.parse_args()
    lib = _capi.lib()
    rng = np.random.default_rng(0)
    D, M = 128, 32
    cw = rng.random((M, 256, D // M), dtype=np.float32)
    codes = rng.integers(0, 256, (a.n, M), dtype=np.uint8)
    e = main.RiiCpp(cw, False, l2_variant=16)
    e.add_codes(codes, False)
    e.reconfigure(a.nlist, 2)
    L = min(a.n, 32 * (a.n // a.nlist))
    Q = rng.random((a.calls, D), dtype=np.float32)
    empty = np.empty(0, np.int64)
    out = {"N": a.n, "nlist": a.nlist, "L": L, "topk": a.topk}
    ids = np.empty(a.topk, np.int64)
    dists = np.empty(a.topk, np.float32)
    pi, pd = ids.ctypes.data_as(C.POINTER(C.c_int64)), dists.ctypes.data_as(C.POINTER(C.c_float))
    for name, fn in (("ivf", lambda q: e.query_ivf(q, a.topk, empty, L)), ("linear", lambda q: e.query_linear(q, a.topk, empty))):
        for q in Q[:20]:
            fn(q)
        t0 = time.perf_counter()
        for q in Q:
            fn(q)
        out["wrapper_%s_us" % name] = round((time.perf_counter() - t0) / len(Q) * 1e6, 2)
    qp = [q.ctypes.data_as(C.POINTER(C.c_float)) for q in Q]
    for name, call in (("ivf", lambda p: lib.rii_query_ivf(e._h, p, a.topk, None, 0, L, pi, pd)),
                       ("linear", lambda p: lib.rii_query_linear(e._h, p, a.topk, None, 0, pi, pd))):
        for zc in (1, 0):
            e.set_option("zero_copy", zc)
            for p in qp[:20]:
                call(p)
            t0 = time.perf_counter()
            for p in qp:
                call(p)
            out["capi_%s_zero_copy%d_us" % (name, zc)] = round((time.perf_counter() - t0) / len(qp) * 1e6, 2)
        e.set_option("zero_copy", 1)
        lib.rii_profile_enable(e._h, 1)
        lib.rii_profile_reset(e._h)
        for p in qp[:200]:
            call(p)
        k = {}
        for kn in ("dtable", "coarse_rank", "scan_ivf", "scan_linear", "merge", "plan"):
            m_, n_ = C.c_double(0), C.c_int64(0)
            lib.rii_profile_get(e._h, kn.encode(), C.byref(m_), C.byref(n_))
            if n_.value:
                k[kn] = {"us_per_call": round(m_.value / 200 * 1e3, 2), "launches_per_call": n_.value / 200}
        out["kernels_%s" % name] = k
        lib.rii_profile_enable(e._h, 0)
    # the linear scan forced onto the streaming engine (auto uses the natural-layout kernels below 2^21 codes)
    e.set_option("scan_kernel", 4)
    call = lambda p: lib.rii_query_linear(e._h, p, a.topk, None, 0, pi, pd)
    for p in qp[:20]:
        call(p)
    t0 = time.perf_counter()
    for p in qp:
        call(p)
    out["capi_linear_stream_engine_us"] = round((time.perf_counter() - t0) / len(qp) * 1e6, 2)
    e.set_option("scan_kernel", 0)
    # per-CTA phase clocks of one single-query IVF call (the fused multi-CTA kernel)
    e.set_option("debug_clocks", 1)
    lib.rii_query_ivf(e._h, qp[0], a.topk, None, 0, L, pi, pd)
    lib.rii_query_ivf(e._h, qp[1], a.topk, None, 0, L, pi, pd)
    parts = max(1, min(148, (L + 12 * 128 - 1) // (12 * 128)))
    clk = np.zeros((parts, 8), np.int64)
    if lib.rii_debug_clocks(e._h, parts, clk.ctypes.data_as(C.POINTER(C.c_int64))) == 0:
        t0 = clk[:, 0]
        ph = {"table_built": clk[:, 4] - t0, "coarse_pass_done": clk[:, 5] - clk[:, 4], "selected": clk[:, 6] - clk[:, 5],
              "planned": clk[:, 1] - clk[:, 6], "scan": clk[:, 2] - clk[:, 1], "tail_merge": clk[:, 3] - clk[:, 2], "total": clk[:, 3] - t0}
        out["ivf_single_query_cta_cycles"] = {k: {"mean": float(v.mean()), "max": float(v.max())} for k, v in ph.items()}
        out["ivf_single_query_ctas"] = parts
    e.set_option("debug_clocks", 0)
    # an empty call for scale: ctypes + one trivial C function
    t0 = time.perf_counter()
    for _ in range(20000):
        lib.rii_get_nlist(e._h)
    out["ctypes_trivial_call_us"] = round((time.perf_counter() - t0) / 20000 * 1e6, 3)
    print(json.dumps(out))


if __name__ == "__main__":
    run()
