"""Kernel-level measurements of the scan kernels on one B200 (not a bench line; feeds profiles/ and DESIGN.md).
Usage: python tools/microbench.py [--n 64000000] [--what linear,ivf,assign]"""
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


def prof(lib, e, name):
    m, n = C.c_double(0), C.c_int64(0)
    lib.rii_profile_get(e._h, name.encode(), C.byref(m), C.byref(n))
    return m.value, n.value


def main_():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64000000)
    ap.add_argument("--m", type=int, default=32)
    ap.add_argument("--what", default="linear,assign")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--scan-kernel", type=int, default=0)
    ap.add_argument("--stream-ctas", type=int, default=0)
    a = ap.parse_args()
    lib = _capi.lib()
    dev = torch.device("cuda", 0)
    M, Ks, D = a.m, 256, 128
    rng = np.random.default_rng(0)
    cw = rng.random((M, Ks, D // M), dtype=np.float32)
    e = main.RiiCpp(cw, False, l2_variant=16)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    N = a.n
    chunk = 8000000
    if "linear" in a.what:
        for s in range(0, N, chunk):
            n = min(chunk, N - s)
            e.add_codes(torch.randint(0, 256, (n, M), dtype=torch.uint8).numpy(), False)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    sp = C.c_void_p(st.cuda_stream)
    lib.rii_profile_enable(e._h, 1)
    e.set_option("scan_kernel", a.scan_kernel)
    e.set_option("stream_ctas", a.stream_ctas)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if "linear" in a.what:
        for B, k in [(1, 1), (1, 10), (1, 100), (4, 10), (16, 10)]:
            Q = torch.rand((B, D), device=dev)
            oi = torch.empty((B, k), dtype=torch.int64, device=dev)
            od = torch.empty((B, k), dtype=torch.float32, device=dev)
            oc = torch.empty((B,), dtype=torch.int32, device=dev)
            for it in range(3):
                _capi.check(lib.rii_query_batch_dev(e._h, C.c_void_p(Q.data_ptr()), B, k, None, 0, 0, 0,
                                                    C.c_void_p(oi.data_ptr()), C.c_void_p(od.data_ptr()),
                                                    C.c_void_p(oc.data_ptr()), sp))
            torch.cuda.synchronize()
            lib.rii_profile_reset(e._h)
            t_tot = 0.0
            for it in range(a.reps):
                flush.zero_()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record(st)
                _capi.check(lib.rii_query_batch_dev(e._h, C.c_void_p(Q.data_ptr()), B, k, None, 0, 0, 0,
                                                    C.c_void_p(oi.data_ptr()), C.c_void_p(od.data_ptr()),
                                                    C.c_void_p(oc.data_ptr()), sp))
                ev1.record(st)
                torch.cuda.synchronize()
                t_tot += ev0.elapsed_time(ev1)
            ms, n = prof(lib, e, "scan_linear")
            per = ms / n
            gbs = B * N * M / (per * 1e-3) / 1e9
            print(json.dumps({"what": "linear", "scan_kernel": a.scan_kernel, "stream_ctas": a.stream_ctas, "N": N, "M": M, "B": B, "topk": k, "scan_ms": round(per, 4),
                              "query_ms": round(t_tot / a.reps, 4), "code_GBps": round(gbs, 1),
                              "frac_hbm_per_query_bytes": round(N * M / (per / B * 1e-3) / 1e9 / peak, 4),
                              "lookups_per_s_T": round(B * N * M / (per * 1e-3) / 1e12, 3)}))
    if "ivf" in a.what:
        # C2-shaped batch (N = 1M random codes, nlist = 1000, L = 32000, topk = 1) with per-CTA phase clocks
        n_iv, B, L = 1000000, 2048, 32000
        e3 = main.RiiCpp(cw, False, l2_variant=16)
        e3.add_codes(torch.randint(0, 256, (n_iv, M), dtype=torch.uint8).numpy(), False)
        e3.reconfigure(1000, 1)
        lib.rii_profile_enable(e3._h, 1)
        e3.set_option("debug_clocks", 1)
        e3.set_option("scan_kernel", a.scan_kernel)
        e3.set_option("stream_ctas", a.stream_ctas)
        Q = torch.rand((B, D), device=dev)
        oi = torch.empty((B, 1), dtype=torch.int64, device=dev)
        od = torch.empty((B, 1), dtype=torch.float32, device=dev)
        oc = torch.empty((B,), dtype=torch.int32, device=dev)
        for fuse in (1, 0):
            e3.set_option("fuse_coarse", fuse)
            for it in range(5):
                if it == 2:
                    torch.cuda.synchronize()
                    lib.rii_profile_reset(e3._h)
                _capi.check(lib.rii_query_batch_dev(e3._h, C.c_void_p(Q.data_ptr()), B, 1, None, 0, L, 1,
                                                    C.c_void_p(oi.data_ptr()), C.c_void_p(od.data_ptr()),
                                                    C.c_void_p(oc.data_ptr()), sp))
            torch.cuda.synchronize()
            clk = np.zeros((B, 8), np.int64)
            _capi.check(lib.rii_debug_clocks(e3._h, B, clk.ctypes.data_as(C.POINTER(C.c_int64))))
            out = {"what": "ivf_batch", "scan_kernel": a.scan_kernel, "stream_ctas": a.stream_ctas, "fuse_coarse": fuse, "B": B, "L": L}
            for name in ("coarse_rank", "scan_ivf", "dtable"):
                ms, n = prof(lib, e3, name)
                out[name + "_ms"] = round(ms / max(n, 1), 4)
            t0 = clk[:, 0]
            out["cta_cycles_mean"] = {"table_built": float((clk[:, 4] - t0).mean()), "ready_to_scan": float((clk[:, 1] - t0).mean()),
                                      "scan": float((clk[:, 2] - clk[:, 1]).mean()), "tail": float((clk[:, 3] - clk[:, 2]).mean()),
                                      "total": float((clk[:, 3] - t0).mean())}
            if fuse:
                out["cta_cycles_mean"].update({"coarse_pass": float((clk[:, 5] - clk[:, 4]).mean()),
                                               "pool_sort": float((clk[:, 6] - clk[:, 5]).mean()),
                                               "plan": float((clk[:, 1] - clk[:, 6]).mean()), "pool_size": float(clk[:, 7].mean())})
            print(json.dumps(out))
        e3.set_option("fuse_coarse", 1)
        e3.set_option("debug_clocks", 0)
        for topk in (10, 100):
            oi2 = torch.empty((B, topk), dtype=torch.int64, device=dev)
            od2 = torch.empty((B, topk), dtype=torch.float32, device=dev)
            for it in range(5):
                if it == 2:
                    torch.cuda.synchronize()
                    lib.rii_profile_reset(e3._h)
                _capi.check(lib.rii_query_batch_dev(e3._h, C.c_void_p(Q.data_ptr()), B, topk, None, 0, L, 1,
                                                    C.c_void_p(oi2.data_ptr()), C.c_void_p(od2.data_ptr()),
                                                    C.c_void_p(oc.data_ptr()), sp))
            torch.cuda.synchronize()
            ms, n = prof(lib, e3, "scan_ivf")
            print(json.dumps({"what": "ivf_batch_topk", "topk": topk, "B": B, "L": L, "scan_ivf_ms": round(ms / max(n, 1), 4)}))
    if "assign" in a.what:
        n_as = min(N, 1000000)
        codes = torch.randint(0, 256, (n_as, M), dtype=torch.uint8).numpy()
        for ak in (1, 0, 3):  # natural-layout k_assign, streaming engine (2 CTAs / SM), streaming engine (1 CTA / SM)
            e2 = main.RiiCpp(cw, False, l2_variant=16)
            e2.set_option("assign_kernel", ak)
            e2.add_codes(codes, False)
            e2.reconfigure(1000, 1)  # warm-up: Dm, skew copies, allocations
            torch.cuda.synchronize()
            lib.rii_profile_enable(e2._h, 1)
            lib.rii_profile_reset(e2._h)
            t0 = time.time()
            e2.reconfigure(1000, 5)
            torch.cuda.synchronize()
            dt = time.time() - t0
            ms, n = prof(lib, e2, "assign")
            print(json.dumps({"what": "reconfigure", "assign_kernel": ak, "N": n_as, "M": M, "nlist": 1000, "iter": 5, "seconds": round(dt, 4),
                              "assign_ms_total": round(ms, 3), "assign_launches": n,
                              "lookups_per_s_T": round((n_as + 5 * 100000) * 1000 * M / (ms * 1e-3) / 1e12, 3)}))
            del e2


if __name__ == "__main__":
    main_()
