#!/usr/bin/env python
"""bench.py -- queries/sec of the ADC hot path on BASELINE.json's metric configuration.

Workload (BASELINE.json configs[1], "C2"): N=1M, D=128, M=32, Ks=256, IVF nlist=1000, L=32*L0=32000
("nprobe=32"), topk=1 (recall@1), synthetic float32 vectors U[0,1)^D (README.md:84-87 of the reference).
A *step* is one batch of B queries through the hot path (distance tables -> coarse ranking -> posting-list
scan -> top-k).

  value : whole-job queries/sec with queries and result buffers already resident in HBM (CUDA events on the
          launching stream, L2 flushed between steps outside the timed events)
  e2e   : the same metric through the reference-facing C ABI call rii_query_batch() with HOST (pinned)
          buffers -- H2D of the queries and D2H of ids/dists/counts inside the timed region
  --impl reference : the UNMODIFIED reference (oracle/_ref/fast_*: its own sources and flags) driven through
          its own single-query API on this box's host cores (one process per core, each a loop of
          main.RiiCpp.query_ivf calls like examples/benchmark/run_sift1m.py:26-30)

Multi-GPU (torchrun, one rank per GPU), total work per step fixed -> "strong":
  default  : N = 1M fits every GPU many times over, so the index is REPLICATED and each rank answers B/G of the step's
             queries and returns their results (independent units: no exchange step, no collective).  The same steps with
             one packed NCCL all-gather (every rank holds the whole batch; --gather) are measured beside the headline.
             The sharded C5 / C4 legs (`sharded_large`) are where the path has real exchange steps.
  --shard  : the index is partitioned by contiguous id range (what C4/C5-sized indexes need): every rank
             scans its shard for every query, per-shard top-k are all-gathered and merged (k_merge_shards).  At
             N = 1M this divides only the scan, not the per-query table / coarse work: measured in profiles/.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(N=1000000, D=128, M=32, Ks=256, nlist=1000, L=32000, topk=1, iter=5)
HBM_FALLBACK_GBS = 6650.0


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ data ----
def make_data(nq, device, clustered=False):
    """Synthetic vectors, a PQ trained on a 20k sample (scipy k-means), codes by exact nearest codeword and the
    exact float-L2 ground truth for recall@1 -- all *setup* (torch on the GPU), outside any timed region.
    clustered: vectors on a 16-dimensional manifold, x = tanh(z W) + 0.02 noise with z ~ N(0, I_16), instead of U[0,1)^D;
    queries from the same distribution -- data with structure (as real descriptors have), on which a 32-byte PQ code
    resolves neighbours and recall@1 means something (uniform 128-d data has no structure to quantise)."""
    import torch
    from rii_b200 import pq
    N, D, M, Ks = CFG["N"], CFG["D"], CFG["M"], CFG["Ks"]
    Ds = D // M
    g = torch.Generator(device=device).manual_seed(123)
    g2 = torch.Generator(device=device).manual_seed(456)
    if clustered:
        W0 = torch.randn((16, D), generator=g, device=device, dtype=torch.float32) / 4.0

        def draw(n, gen):
            z = torch.randn((n, 16), generator=gen, device=device, dtype=torch.float32)
            return torch.tanh(z @ W0) + 0.02 * torch.randn((n, D), generator=gen, device=device, dtype=torch.float32)
        X, Q = draw(N, g), draw(nq, g2)
    else:
        X = torch.rand((N, D), generator=g, device=device, dtype=torch.float32)
        Q = torch.rand((nq, D), generator=g2, device=device, dtype=torch.float32)
    codec = pq.PQ(M=M, Ks=Ks, verbose=False).fit(X[:20000].cpu().numpy(), iter=10, seed=123)
    cw = torch.from_numpy(codec.codewords).to(device)
    codes = torch.empty((N, M), dtype=torch.uint8, device=device)
    for m in range(M):
        sub = X[:, m * Ds:(m + 1) * Ds]
        codes[:, m] = torch.cdist(sub, cw[m]).argmin(1).to(torch.uint8)
    gt = torch.empty(nq, dtype=torch.int64, device=device)
    xn = (X * X).sum(1)
    for s in range(0, nq, 1024):
        q = Q[s:s + 1024]
        gt[s:s + 1024] = (xn[None, :] - 2.0 * q @ X.T).argmin(1)
    del X, xn
    torch.cuda.empty_cache()
    return codec.codewords, codes.cpu().numpy(), Q.cpu().numpy(), gt.cpu().numpy()


def make_data_cpu(nq):
    """Same distribution without a GPU (reference arm on a box whose GPU we do not touch): numpy + scipy."""
    from rii_b200 import pq
    N, D, M, Ks = CFG["N"], CFG["D"], CFG["M"], CFG["Ks"]
    rng = np.random.default_rng(123)
    X = rng.random((N, D), dtype=np.float32)
    Q = np.random.default_rng(456).random((nq, D), dtype=np.float32)
    codec = pq.PQ(M=M, Ks=Ks, verbose=False).fit(X[:20000], iter=10, seed=123)
    codes = codec.encode(X)
    return codec.codewords, codes, Q, None


# ---------------------------------------------------------------------------------------- clocks ----
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def _run_nvml(self):
        """In-process NVML sampling every ~5 ms (an nvidia-smi call takes longer than a whole timed region)."""
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = [("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap)]
        while not self.stop_flag:
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.samples.append([str(sm), str(mx)] + ["Active" if r & b else "Not Active" for _, b in bits])
            time.sleep(0.005)

    def run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass  # no NVML binding / call failed: fall back to polling nvidia-smi
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                               "--format=csv,noheader,nounits"], timeout=5).decode().strip()
                self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ ours ----
def _ptr(t):
    return C.c_void_p(t.data_ptr())


def kernel_ms(lib, e, name):
    m_, n_ = C.c_double(0), C.c_int64(0)
    lib.rii_profile_get(e._h, name.encode(), C.byref(m_), C.byref(n_))
    return m_.value, n_.value


class PackedOut(object):
    """One byte buffer [ids int64 (B, k) | dists float32 (B, k) | counts int32 (B)] per rank: the library writes its three
    outputs straight into it and ONE all-gather moves everything (three NCCL launches per step were 9 % of it)."""

    def __init__(self, torch, B, k, dev, world):
        self.B, self.k = B, k
        self.nbytes = (B * k * 12 + B * 4 + 7) // 8 * 8
        self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=dev)
        self.ids = self.buf[:B * k * 8].view(torch.int64).view(B, k)
        self.d = self.buf[B * k * 8:B * k * 12].view(torch.float32).view(B, k)
        self.c = self.buf[B * k * 12:B * k * 12 + B * 4].view(torch.int32)
        self.gathered = torch.zeros((world, self.nbytes), dtype=torch.uint8, device=dev) if world > 1 else None

    def rank_ids(self, torch, r):
        return self.gathered[r, :self.B * self.k * 8].view(torch.int64).view(self.B, self.k)


def large_sharded_leg(torch, dist, lib, main, _capi, rank, world, dev, st, sp, peak, cfg, steps, warmup, flush):
    """BASELINE configs C4 / C5 at their real per-GPU size (weak scaling: every GPU holds Nl codes, N_total = Nl * G; at
    G = 8 this IS C4 / C5).  Codes are drawn as randint (throughput-only runs); list assignment is drawn too
    -- K6 over 65536 centers is an index-build cost (tools/build_large.py measures the real build), the scan does not care
    which rows sit in a list.  Every rank ranks the lists for its B / G queries, the rankings are all-gathered, every rank
    scans its shard for all B queries, per-shard top-k are all-gathered (one packed collective) and merged."""
    D, M, Ks, nlist, Nl, B, k = cfg["D"], cfg["M"], 256, cfg["nlist"], cfg["Nl"], cfg["B"], cfg["topk"]
    N_total = Nl * world
    L0 = int(round(N_total / nlist))
    L = min(32 * L0, N_total)
    rng = np.random.default_rng(2024)
    cw = rng.random((M, Ks, D // M), dtype=np.float32)
    centers = rng.integers(0, Ks, (nlist, M), dtype=np.uint8)
    Q = torch.from_numpy(rng.random((4 * B, D), dtype=np.float32)).to(dev)
    e = main.RiiCpp(cw, False, device=dev.index, l2_variant=16)
    t0 = time.time()
    _capi.check(lib.rii_reserve(e._h, Nl))
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    keep = []  # (first row, codes) of the chunks kept for the parity check
    chunk = 1 << 23
    for s0 in range(0, Nl, chunk):
        c = min(chunk, Nl - s0)
        t = torch.randint(0, 256, (c, M), dtype=torch.uint8, device=dev, generator=gen)
        _capi.check(lib.rii_add_codes_dev(e._h, _ptr(t), c, 0))
        if s0 == 0:
            keep.append((0, t))
    assign = torch.randint(0, nlist, (Nl,), dtype=torch.int32, device=dev, generator=gen)
    _capi.check(lib.rii_set_shard(e._h, rank * Nl, N_total))
    _capi.check(lib.rii_set_lists_dev(e._h, centers.ctypes.data_as(C.POINTER(C.c_uint8)), nlist, _ptr(assign)))
    lens = torch.bincount(assign, minlength=nlist).to(torch.int32)
    if world > 1:
        g = torch.empty((world, nlist), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(g.view(-1), lens)
    else:
        g = lens[None]
    glob = g.sum(0, dtype=torch.int32).cpu().numpy()
    pre = (g[:rank].sum(0, dtype=torch.int32) if rank else torch.zeros_like(lens)).cpu().numpy()
    _capi.check(lib.rii_set_global_lengths(e._h, glob.ctypes.data_as(C.POINTER(C.c_int32)), pre.ctypes.data_as(C.POINTER(C.c_int32))))
    torch.cuda.synchronize()
    t_build = time.time() - t0
    if os.environ.get("RII_LARGE_CTAS"):
        lib.rii_set_option(e._h, b"stream_ctas", int(os.environ["RII_LARGE_CTAS"]))
    w = _capi.check(lib.rii_coarse_width(e._h, L))
    Bl = B // world
    ranked_l = torch.empty((Bl, w), dtype=torch.int32, device=dev)
    ranked = torch.empty((B, w), dtype=torch.int32, device=dev) if world > 1 else ranked_l
    po = PackedOut(torch, B, k, dev, world)
    flags = torch.zeros((B,), dtype=torch.int32, device=dev)
    f_ids = torch.empty((B, k), dtype=torch.int64, device=dev)
    f_d = torch.empty((B, k), dtype=torch.float32, device=dev)
    f_c = torch.empty((B,), dtype=torch.int32, device=dev)

    def step(i):
        q = Q[(i % 4) * B:(i % 4 + 1) * B]
        _capi.check(lib.rii_coarse_rank_dev(e._h, _ptr(q[rank * Bl:(rank + 1) * Bl]), Bl, k, L, _ptr(ranked_l), sp))
        if world > 1:
            dist.all_gather_into_tensor(ranked.view(-1), ranked_l.view(-1))
        _capi.check(lib.rii_query_ranked_dev(e._h, _ptr(q), B, k, L, _ptr(ranked), _ptr(po.ids), _ptr(po.d), _ptr(po.c), _ptr(flags), sp))
        if world > 1:
            dist.all_gather_into_tensor(po.gathered.view(-1), po.buf)
            _capi.check(lib.rii_merge_shards_packed_dev(e._h, _ptr(po.gathered), po.nbytes, world, B, k, _ptr(f_ids), _ptr(f_d), _ptr(f_c), sp))
            return f_ids
        return po.ids

    for i in range(warmup):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n_flagged = int((flags & 1).sum().item())
    lib.rii_profile_enable(e._h, 1)
    lib.rii_profile_reset(e._h)
    evs = []
    for i in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        step(warmup + i)
        b.record(st)
        evs.append((a, b))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    scan_ms, scan_n = kernel_ms(lib, e, "scan_ivf")
    coarse_ms, coarse_n = kernel_ms(lib, e, "coarse_rank")
    lib.rii_profile_enable(e._h, 0)
    per_scan = scan_ms / max(scan_n, 1)
    # what this rank's scan launch actually streamed: the planned local candidates (ids are read for survivors only)
    local_rows = float(L) / world
    alg = B * (local_rows * M + 4 * M * Ks)
    out = {"workload": "%s: N=%d (%d per GPU x %d) D=%d M=%d nlist=%d IVF L=%d (w=%d lists) topk=%d batch=%d; random codes + random list "
                       "assignment (throughput only)" % (cfg["name"], N_total, Nl, world, D, M, nlist, L, w, k, B),
           "ms_per_batch_this_rank": round(ms / steps, 4), "index_build_s": round(t_build, 2), "flagged_queries": n_flagged,
           "scan_kernel_ms": round(per_scan, 4), "coarse_kernel_ms": round(coarse_ms / max(coarse_n, 1), 4),
           "scan_algorithmic_bytes": int(alg), "scan_GBps_per_gpu": round(alg / (per_scan * 1e-3) / 1e9, 1) if per_scan > 0 else None,
           "scan_frac_of_hbm_peak": round(alg / (per_scan * 1e-3) / 1e9 / peak, 4) if per_scan > 0 else None}
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["ms_per_batch"] = round(float(t.item()) / steps, 4)
    out["queries_per_s"] = round(steps * B / (float(t.item()) * 1e-3), 1)
    # ---- parity at full scale (rank 0): the shard's part of the answer against a numpy / oracle restatement --------------
    step(0)  # (collectives inside: every rank runs it; only rank 0 checks its shard's part below)
    torch.cuda.synchronize()
    if rank == 0:
        try:
            from oracle import oracle as O
            q = Q[:B]
            got_ids, got_d, got_c = po.ids.cpu().numpy(), po.d.cpu().numpy(), po.c.cpu().numpy()
            rk = ranked.cpu().numpy()
            ok, why = True, []
            for bq in (0, B // 2, B - 1):
                T = O.dtable(q[bq].cpu().numpy(), cw, 16)
                cd = O.adist_all(T, centers)
                order = np.lexsort((np.arange(nlist), cd))[:w]
                if not np.array_equal(order, rk[bq]):
                    ok = False
                    why.append("ranking of query %d" % bq)
                P, rows = 0, []
                for j, no in enumerate(order):  # SURVEY A.3 with global lengths, this shard's slice of every list
                    f = int(glob[no])
                    take = min(f, L - P)
                    P += take
                    mine = torch.nonzero(assign == int(no)).flatten()
                    lt = int(np.clip(take - int(pre[no]), 0, mine.numel()))
                    rows.append(mine[:lt])
                    if P >= L or (j == w - 1 and P >= k):
                        break
                rows = torch.cat(rows)
                # rows of the first chunk only are still at hand on the device: regenerate the others
                codes_rows = regen_rows(torch, dev, rank, Nl, M, chunk, rows)
                dd = O.adist_all(T, codes_rows)
                o = np.lexsort((rows.cpu().numpy(), dd))[:k]
                exp_ids = rows.cpu().numpy()[o] + rank * Nl
                if not (int(got_c[bq]) == len(o) and np.array_equal(got_ids[bq, :len(o)], exp_ids) and
                        np.array_equal(got_d[bq, :len(o)].view(np.uint32), dd[o].view(np.uint32))):
                    ok = False
                    why.append("ivf result of query %d: got %s %s, expected %s %s (%d candidate rows)" % (
                        bq, got_ids[bq, :int(got_c[bq])].tolist(), got_d[bq, :int(got_c[bq])].tolist(), exp_ids.tolist(), dd[o].tolist(), len(rows)))
            # sampled-shard LINEAR parity: the first 200 000 local rows as target ids == oracle scan of those rows
            ns = min(200000, Nl)
            tids = torch.arange(rank * Nl, rank * Nl + ns, dtype=torch.int64, device=dev)
            li = torch.empty((1, 5), dtype=torch.int64, device=dev)
            ld = torch.empty((1, 5), dtype=torch.float32, device=dev)
            lc = torch.empty((1,), dtype=torch.int32, device=dev)
            lib.rii_set_option(e._h, b"scan_kernel", 4)
            _capi.check(lib.rii_query_batch_dev(e._h, _ptr(q[:1]), 1, 5, _ptr(tids), ns, 0, 0, _ptr(li), _ptr(ld), _ptr(lc), sp))
            lib.rii_set_option(e._h, b"scan_kernel", 0)
            torch.cuda.synchronize()
            exp = O.query_linear(O.dtable(q[0].cpu().numpy(), cw, 16), keep[0][1][:ns].cpu().numpy(), 5)
            if not (np.array_equal(li.cpu().numpy()[0], exp[0] + rank * Nl) and
                    np.array_equal(ld.cpu().numpy()[0].view(np.uint32), exp[1].view(np.uint32))):
                ok = False
                why.append("sampled linear scan: got %s, expected %s" % (li.cpu().numpy()[0].tolist(), (exp[0] + rank * Nl).tolist()))
            out["parity_vs_oracle_at_full_scale"] = "ok (3 IVF queries on this shard: ranking, candidate set, ids and distance bits; " \
                                                    "linear scan over 200000 sampled rows)" if ok else "MISMATCH: " + "; ".join(why)
        except Exception as ex:
            out["parity_vs_oracle_at_full_scale"] = "error: " + repr(ex)
    del e
    torch.cuda.empty_cache()
    return out


def regen_rows(torch, dev, rank, Nl, M, chunk, rows):
    """Codes of the given local rows, regenerated chunk by chunk with the generator large_sharded_leg used."""
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    out = torch.empty((rows.numel(), M), dtype=torch.uint8, device=dev)
    for s0 in range(0, Nl, chunk):
        c = min(chunk, Nl - s0)
        t = torch.randint(0, 256, (c, M), dtype=torch.uint8, device=dev, generator=gen)
        sel = torch.nonzero((rows >= s0) & (rows < s0 + c)).flatten()
        if sel.numel():
            out[sel] = t[rows[sel] - s0]
    return out.cpu().numpy()


def run_ours(args):
    import torch
    import torch.distributed as dist
    from rii_b200 import _capi, main
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))  # a lost rank must not hang the box
    lib = _capi.lib()
    B, K, W = args.batch, args.steps, args.warmup
    nq = B * min(4, K + W)  # a few distinct query batches, cycled
    cw, codes, Q, gt = make_data(nq, dev)
    N, M, D = CFG["N"], CFG["M"], CFG["D"]
    peak, peak_src = peaks()
    props = torch.cuda.get_device_properties(dev)

    # ---- index: replicated per rank (the metric's N = 1M fits every GPU many times over); --shard: id-range shards ------
    e = main.RiiCpp(cw, False, device=local, l2_variant=16)
    t_build = time.time()
    shard = bool(args.shard) and world > 1
    if not shard:
        e.add_codes(codes, False)
        e.reconfigure(CFG["nlist"], CFG["iter"])
    else:
        from rii_b200 import sharded
        sharded.build_shard(e, codes, CFG["nlist"], CFG["iter"], rank, world)
    torch.cuda.synchronize()
    t_build = time.time() - t_build

    st = torch.cuda.Stream(device=dev)  # a real (non-default) stream: events, kernels and NCCL all on it
    torch.cuda.set_stream(st)
    sp = C.c_void_p(st.cuda_stream)
    k, L = CFG["topk"], CFG["L"]
    dQ = torch.from_numpy(Q).to(dev)
    assert B % world == 0
    Bl = B if (world == 1 or shard) else B // world  # queries this rank answers per step
    po = PackedOut(torch, Bl, k, dev, world)
    f_ids = torch.empty((B, k), dtype=torch.int64, device=dev)
    f_d = torch.empty((B, k), dtype=torch.float32, device=dev)
    f_c = torch.empty((B,), dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # Replicas (the default for the metric's N = 1M): every rank answers ITS B / world queries and keeps the results -- the
    # queries are independent units, there is no exchange step on this path, so no collective is timed (--gather adds the
    # packed all-gather that makes every rank hold the whole batch; it is also measured once, beside the headline).
    # Shards: the per-shard top-k MUST be exchanged: one packed all-gather + merge.
    def finish_step(gather):
        """After this rank's results are in po.buf: the collective of the step (if any), and what the caller reads."""
        if world == 1 or (not shard and not gather):
            return po.ids
        dist.all_gather_into_tensor(po.gathered.view(-1), po.buf)
        if not shard:  # replicas: rank r answered queries [r * Bl, (r + 1) * Bl): every rank now holds all of them
            return po.ids
        _capi.check(lib.rii_merge_shards_packed_dev(e._h, _ptr(po.gathered), po.nbytes, world, B, k, _ptr(f_ids), _ptr(f_d), _ptr(f_c), sp))
        return f_ids

    def step_dev(i, gather=None):
        q = dQ[(i % (nq // B)) * B:(i % (nq // B) + 1) * B]
        if world > 1 and not shard:
            q = q[rank * Bl:(rank + 1) * Bl]
        _capi.check(lib.rii_query_batch_dev(e._h, _ptr(q), Bl, k, None, 0, L, 1, _ptr(po.ids), _ptr(po.d), _ptr(po.c), sp))
        return finish_step(bool(args.gather) if gather is None else gather)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm (value) --------------------------------------------------------------
    for i in range(W):
        step_dev(i)
    barrier()
    lib.rii_profile_enable(e._h, 1)
    lib.rii_profile_reset(e._h)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.rii_launch_count()
    evs = []
    got = []
    for i in range(K):
        flush.zero_()  # L2 flush, outside the timed events
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        ret = step_dev(W + i)
        b.record(st)
        evs.append((a, b))
        if i < nq // B:
            got.append((W + i, ret.clone()))  # this rank's answers (replicas: its slice of the step; shards / 1 GPU: the whole step)
    barrier()
    launches = lib.rii_launch_count() - launches0
    sampler.stop_flag = True
    ms = sum(a.elapsed_time(b) for a, b in evs)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # replicas: the same K steps once more WITH (default run) / WITHOUT (--gather run) the packed all-gather of the results
    no_gather = None
    if world > 1 and not shard:
        other = not bool(args.gather)
        evs2 = []
        for i in range(K):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            step_dev(W + i, gather=other)
            b.record(st)
            evs2.append((a, b))
        barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs2)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        no_gather = {"all_gather": other, "queries_per_s": round(K * B / (float(t.item()) * 1e-3), 1), "ms_per_step": round(float(t.item()) / K, 4),
                     "note": ("the same steps with one packed NCCL all-gather (ids | dists | counts) after the search, so that every rank holds "
                              "the whole batch's results" if other else "the same steps without the all-gather") + "; max over ranks"}
    scan_ms, scan_n = kernel_ms(lib, e, "scan_ivf")
    scan_n = scan_n if no_gather is None else scan_n // 2  # (the profile saw both loops; per-launch time is the same)
    scan_ms = scan_ms if no_gather is None else scan_ms / 2
    prof = {}
    for name in ("dtable", "coarse_rank", "scan_ivf", "merge"):
        m_, n_ = kernel_ms(lib, e, name)
        prof[name] = {"ms_total": round(m_, 4), "launches": n_}
    lib.rii_profile_enable(e._h, 0)
    sampler.join(timeout=2)

    # recall@1 (examples/benchmark/util.py:35-58) of what the timed steps returned
    hit = tot = 0
    for i, ids in got:
        s = (i % (nq // B)) * B + (rank * Bl if (world > 1 and not shard) else 0)
        n_ = ids.shape[0]
        hit += int((ids[:, 0].cpu().numpy() == gt[s:s + n_]).sum())
        tot += n_
    if world > 1 and not shard:
        t = torch.tensor([hit, tot], device=dev, dtype=torch.float64)
        dist.all_reduce(t)
        hit, tot = float(t[0].item()), float(t[1].item())
    recall = hit / max(tot, 1)

    # ---- end-to-end arm: host (pinned) buffers.  1 GPU: the reference-facing C-ABI call rii_query_batch (H2D of the queries
    # and D2H of ids / dists / counts inside the call).  N GPUs: every rank copies ITS queries in, answers them, the results
    # are all-gathered on the device and every rank copies the whole batch's results out.
    hQ = torch.from_numpy(Q).pin_memory()
    if world == 1:
        h_ids = torch.empty((B, k), dtype=torch.int64).pin_memory()
        h_d = torch.empty((B, k), dtype=torch.float32).pin_memory()
        h_c = torch.empty((B,), dtype=torch.int32).pin_memory()

        def step_host(i):
            q = hQ[(i % (nq // B)) * B:(i % (nq // B) + 1) * B]
            _capi.check(lib.rii_query_batch(e._h, C.cast(q.data_ptr(), C.POINTER(C.c_float)), B, k, None, 0, L, 1,
                                            C.cast(h_ids.data_ptr(), C.POINTER(C.c_int64)),
                                            C.cast(h_d.data_ptr(), C.POINTER(C.c_float)),
                                            C.cast(h_c.data_ptr(), C.POINTER(C.c_int32))))
        h2d, d2h = B * D * 4, B * k * 12 + B * 4
    else:
        dq_l = torch.empty((Bl, D), dtype=torch.float32, device=dev)
        h_all = torch.empty((world, po.nbytes), dtype=torch.uint8).pin_memory()
        h_own = torch.empty((po.nbytes,), dtype=torch.uint8).pin_memory()
        h_fin = torch.empty((B * k * 12 + B * 4,), dtype=torch.uint8).pin_memory()

        hl_ids = torch.empty((Bl, k), dtype=torch.int64).pin_memory()
        hl_d = torch.empty((Bl, k), dtype=torch.float32).pin_memory()
        hl_c = torch.empty((Bl,), dtype=torch.int32).pin_memory()

        def step_host(i):
            q = hQ[(i % (nq // B)) * B:(i % (nq // B) + 1) * B]
            if not shard:
                q = q[rank * Bl:(rank + 1) * Bl]
            if not shard and not args.gather:
                # replicas without a collective: exactly the 1-GPU end-to-end path, per rank -- the reference-facing C-ABI call with
                # host buffers (H2D of this rank's queries, chunk-pipelined inside the call, and D2H of their results)
                _capi.check(lib.rii_query_batch(e._h, C.cast(q.data_ptr(), C.POINTER(C.c_float)), Bl, k, None, 0, L, 1,
                                                C.cast(hl_ids.data_ptr(), C.POINTER(C.c_int64)),
                                                C.cast(hl_d.data_ptr(), C.POINTER(C.c_float)),
                                                C.cast(hl_c.data_ptr(), C.POINTER(C.c_int32))))
                return
            dq_l.copy_(q, non_blocking=True)
            _capi.check(lib.rii_query_batch_dev(e._h, _ptr(dq_l), Bl, k, None, 0, L, 1, _ptr(po.ids), _ptr(po.d), _ptr(po.c), sp))
            ret = finish_step(bool(args.gather))
            if not shard and args.gather:
                h_all.copy_(po.gathered, non_blocking=True)
            elif not shard:
                h_own.copy_(po.buf, non_blocking=True)  # this rank's results
            else:
                h_fin[:B * k * 8].copy_(f_ids.view(-1).view(torch.uint8), non_blocking=True)
                h_fin[B * k * 8:B * k * 12].copy_(f_d.view(-1).view(torch.uint8), non_blocking=True)
                h_fin[B * k * 12:].copy_(f_c.view(torch.uint8), non_blocking=True)
            st.synchronize()
        h2d, d2h = Bl * D * 4, ((world * po.nbytes if args.gather else Bl * k * 12 + Bl * 4) if not shard else B * k * 12 + B * 4)
    for i in range(W):
        step_host(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        step_host(W + i)  # synchronous: returns after the D2H of the results
    barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = {"value": round(K * B / dt, 1), "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "note": ("per rank; rii_query_batch (C ABI, host buffers) on every rank" if (world > 1 and not shard and not args.gather) else
                    "per rank") if world > 1 else "rii_query_batch (C ABI, host buffers)"}
    if world == 1:
        # single-query latency through the reference's own call shape (query_ivf, one query per call)
        qs = Q[:200]
        for q in qs[:10]:
            e.query_ivf(q, k, np.empty(0, np.int64), L)
        t1 = time.perf_counter()
        for q in qs:
            e.query_ivf(q, k, np.empty(0, np.int64), L)
        e2e["single_query_call_us"] = round((time.perf_counter() - t1) / len(qs) * 1e6, 1)

    # ---- C3 (BASELINE configs[2]): subset search, target_ids = 100k random ids, 256 queries per call -----------------
    subset = None
    if world == 1 and not args.quick:
        try:
            Bs, S = 256, 100000
            tids = torch.from_numpy(np.sort(np.random.default_rng(3).choice(N, S, replace=False)).astype(np.int64)).to(dev)
            so = PackedOut(torch, Bs, k, dev, 1)
            subset = {"workload": "C3: N=1M M=32 target_ids=100k random (sorted), %d queries per call, topk=1" % Bs}
            for name, method, Ls in (("linear", 0, 0), ("ivf", 1, L)):
                ev = []
                lib.rii_profile_enable(e._h, 1)
                lib.rii_profile_reset(e._h)
                for it in range(13):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(st)
                    _capi.check(lib.rii_query_batch_dev(e._h, _ptr(dQ[it * Bs:(it + 1) * Bs]), Bs, k, _ptr(tids), S, Ls, method,
                                                        _ptr(so.ids), _ptr(so.d), _ptr(so.c), sp))
                    b.record(st)
                    if it >= 3:
                        ev.append((a, b))
                torch.cuda.synchronize()
                subset[name + "_queries_per_s"] = round(len(ev) * Bs / (sum(a.elapsed_time(b) for a, b in ev) * 1e-3), 1)
                subset[name + "_ms_per_call"] = round(sum(a.elapsed_time(b) for a, b in ev) / len(ev), 4)
                subset[name + "_kernel_ms_per_call"] = {kn: round(kernel_ms(lib, e, kn)[0] / 13, 4) for kn in
                                                        ("subset_build", "scan_linear", "scan_ivf", "coarse_rank", "merge", "sort")}
                lib.rii_profile_enable(e._h, 0)
            # the same IVF subset search with the sub-index prepared once per target set (rii_subset_begin_dev) and reused
            # by every batch (rii_subset_query_dev): what a caller does who asks many queries against one subset
            cnts = torch.empty((CFG["nlist"],), dtype=torch.int32, device=dev)
            _capi.check(lib.rii_subset_begin_dev(e._h, _ptr(tids), S, _ptr(cnts), sp))
            ev = []
            for it in range(13):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st)
                _capi.check(lib.rii_subset_query_dev(e._h, _ptr(dQ[it * Bs:(it + 1) * Bs]), Bs, k, L, _ptr(so.ids), _ptr(so.d), _ptr(so.c), sp))
                b.record(st)
                if it >= 3:
                    ev.append((a, b))
            torch.cuda.synchronize()
            subset["ivf_prepared_subset_queries_per_s"] = round(len(ev) * Bs / (sum(a.elapsed_time(b) for a, b in ev) * 1e-3), 1)
        except Exception as ex:
            subset = {"error": repr(ex)}

    # ---- the same C2 shape on CLUSTERED synthetic data: recall@1 that says something, and the rate on non-uniform lists ----
    clustered = None
    if world == 1 and not args.quick:
        try:
            Bc = 8192
            cw2, codes2, Q2, gt2 = make_data(Bc, dev, clustered=True)
            ec = main.RiiCpp(cw2, False, device=local, l2_variant=16)
            ec.add_codes(codes2, False)
            ec.reconfigure(CFG["nlist"], CFG["iter"])
            dQ2 = torch.from_numpy(Q2).to(dev)
            co = PackedOut(torch, Bc, k, dev, 1)
            ev = []
            for it in range(6):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st)
                _capi.check(lib.rii_query_batch_dev(ec._h, _ptr(dQ2), Bc, k, None, 0, L, 1, _ptr(co.ids), _ptr(co.d), _ptr(co.c), sp))
                b.record(st)
                if it >= 3:
                    ev.append((a, b))
            torch.cuda.synchronize()
            lens = np.diff(ec.posting_lists_csr()[0])
            clustered = {"data": "x = tanh(z W) + 0.02 noise, z ~ N(0, I_16): N=1M vectors on a 16-d manifold in R^128, same PQ / IVF shape as C2",
                         "recall_at_1": round(float((co.ids[:, 0].cpu().numpy() == gt2).mean()), 4),
                         "queries_per_s": round(len(ev) * Bc / (sum(a.elapsed_time(b) for a, b in ev) * 1e-3), 1),
                         "posting_list_length_min_mean_max": [int(lens.min()), float(lens.mean()), int(lens.max())]}
            del ec
            torch.cuda.empty_cache()
        except Exception as ex:
            clustered = {"error": repr(ex)}

    # ---- the HBM-bound side of the same engine: linear PQ-code scan over N >> L2 (north_star's ">= 70 % of the HBM
    # roofline on the code scan at N = 1B"); random codes (throughput only), 1 query per launch, CUDA events inside the library
    lin = None
    if world == 1 and args.linear_n > 0:
        for nl in ([int(args.linear_n), 64000000] if int(args.linear_n) > 64000000 else [int(args.linear_n)]):
            el = None
            try:
                el = main.RiiCpp(cw, False, device=local, l2_variant=16)
                _capi.check(lib.rii_reserve(el._h, nl))
                chunk = 1 << 24
                gen = torch.Generator(device=dev).manual_seed(7)
                for s0 in range(0, nl, chunk):
                    c = min(chunk, nl - s0)
                    t = torch.randint(0, 256, (c, M), dtype=torch.uint8, device=dev, generator=gen)
                    _capi.check(lib.rii_add_codes_dev(el._h, _ptr(t), c, 0))
                del t
                lib.rii_profile_enable(el._h, 1)
                q1 = dQ[:1]
                lo = PackedOut(torch, 1, 1, dev, 1)
                for it in range(13):
                    if it == 3:
                        torch.cuda.synchronize()
                        lib.rii_profile_reset(el._h)
                    flush.zero_()
                    _capi.check(lib.rii_query_batch_dev(el._h, _ptr(q1), 1, 1, None, 0, 0, 0, _ptr(lo.ids), _ptr(lo.d), _ptr(lo.c), sp))
                torch.cuda.synchronize()
                m_, n_ = kernel_ms(lib, el, "scan_linear")
                lms = m_ / max(n_, 1)
                lin = {"kernel": "k_scan_stream32<NW=12, linear, 4-stage rings, 1 CTA/SM>",
                       "workload": "linear scan, N=%d M=32 (random codes), topk=1, 1 query/launch" % nl,
                       "bound": "hbm", "launch_ms": round(lms, 4), "algorithmic_bytes_per_launch": nl * M + 4 * M * CFG["Ks"],
                       "achieved": round((nl * M + 4 * M * CFG["Ks"]) / (lms * 1e-3) / 1e9, 1), "unit": "GB/s", "peak": peak}
                lin["frac"] = round(lin["achieved"] / peak, 4)
                del el
                torch.cuda.empty_cache()
                break
            except Exception as ex:  # never lose the headline line over the side measurement (e.g. not enough memory for N = 1B)
                lin = {"error": repr(ex), "N": nl}
                del el
                torch.cuda.empty_cache()

    # ---- C4 / C5 at their real per-GPU size, sharded (weak scaling) ------------------------------------------------------
    large = None
    if not args.no_large:
        large = []
        for cfg in (dict(name="C5", D=96, M=32, nlist=65536, Nl=125000000, B=1024, topk=1),
                    dict(name="C4", D=128, M=64, nlist=10000, Nl=12500000, B=1024, topk=1)):
            try:
                large.append(large_sharded_leg(torch, dist, lib, main, _capi, rank, world, dev, st, sp, peak, cfg, max(3, min(K, 10)), 3, flush))
            except Exception as ex:
                large.append({"workload": cfg["name"], "error": repr(ex)})

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # The dominant kernel (posting-list scan, fused with table build + coarse pass + plan) per launch.  SURVEY 8d's per-query
    # algorithmic bytes: C*M code bytes (C = L candidates) + 4*M*Ks (its distance table); the V*4 bytes of visited ids are
    # not counted (the kernel streams a list-ordered skew64 copy and reads ids for survivors only), nor the nlist*M center bytes.
    # At N = 1M the 32 MB code table is L2-resident, so the binding resource is the shared-memory LOOKUP rate (one 4-byte
    # lookup per code byte; peak = SMs x 32 lookups / clk), not HBM: the roofline is quoted against that, HBM and L2 beside it.
    frac_scanned = 1.0 / world if shard else 1.0
    q_per_launch = K * Bl / max(scan_n, 1)  # the library processes a step in chunks of <= 32768 queries
    alg = q_per_launch * (L * frac_scanned * M + 4 * M * CFG["Ks"])
    lookups = q_per_launch * (L * frac_scanned + CFG["nlist"]) * M
    launch_ms = scan_ms / max(scan_n, 1)
    clocks = sampler.summary()
    sm_mhz = clocks.get("sm_max_mhz") or 1965
    lookup_peak = props.multi_processor_count * 32 * sm_mhz * 1e6 / 1e12
    achieved_l = lookups / (launch_ms * 1e-3) / 1e12 if launch_ms > 0 else None
    achieved_b = alg / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else None
    line = {
        "metric": "queries/sec at recall@1 (N=1M, D=128, M=32)", "value": round(K * B / (ms * 1e-3), 1),
        "unit": "queries/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(ms / K, 4),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic U[0,1)^128 float32 vectors, PQ trained on a 20k sample; ground truth exact L2",
        "config": {"workload": "C2: N=1M D=128 M=32 Ks=256 IVF nlist=1000 L=32000 (w=35 lists) topk=1",
                   "arithmetic": "uint8 codes index float32 tables; float32 adds in the reference's order (bit-exact)",
                   "batch_queries_per_step": B, "l2": "flushed between steps (256 MB write)",
                   "parallelism": "1 GPU" if world == 1 else
                   ("id-range shards x%d + one packed NCCL all-gather of per-shard top-k + merge" % world if shard else
                    "index replicated x%d, queries split%s" % (world, ", one packed NCCL all-gather of the results" if args.gather else
                                                                  "; no collective on this path (each rank returns the results of its own queries)")),
                   "index_build_s": round(t_build, 2)},
        "recall_at_1": round(recall, 4),
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"kernel": "k_scan_persist32<topk=1> (persistent, warp-specialised: 11 scanning warps + 1 producer warp: table, coarse pass, selection, plan, merge)",
                     "bound": "smem-lookup", "achieved": None if achieved_l is None else round(achieved_l, 3), "peak": round(lookup_peak, 3),
                     "unit": "Tlookup/s", "frac": None if achieved_l is None else round(achieved_l / lookup_peak, 4),
                     "peak_source": "%d SMs x 32 four-byte shared-memory lookups / clk x %d MHz" % (props.multi_processor_count, sm_mhz),
                     "traffic": None, "lookups_per_launch": int(lookups), "algorithmic_bytes_per_launch": int(alg),
                     "queries_per_launch": int(q_per_launch), "launch_ms": round(launch_ms, 4),
                     "hbm": {"achieved": None if achieved_b is None else round(achieved_b, 1), "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                             "frac": None if achieved_b is None else round(achieved_b / peak, 4),
                             "note": "algorithmic code + table bytes over the launch time; the 32 MB table is L2-resident at N=1M, so "
                                     "DRAM traffic (`traffic`) is a small fraction of it: the HBM-bound case of the same engine is "
                                     "roofline_linear_scan and the sharded C4/C5 legs"}},
        "kernel_ms": prof,
    }
    if no_gather is not None:
        line["replicas_with_all_gather" if no_gather["all_gather"] else "replicas_without_all_gather"] = no_gather
    if lin is not None:
        line["roofline_linear_scan"] = lin
    if subset is not None:
        line["subset_search"] = subset
    if clustered is not None:
        line["structured_data"] = clustered
    if large is not None:
        line["sharded_large"] = large
    if args.cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_sample(cw, codes, Q)
    try:  # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture (tools/ncu_summary.py)
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        # (captured on a launch of 8192 queries; DRAM traffic of this L2-resident workload is per query: scale to the
        # queries one launch of this run processes)
        t_ivf, q_ivf = tr.get("k_scan_persist32_bytes_per_launch"), tr.get("k_scan_persist32_queries_per_launch", 8192)
        line["roofline"]["traffic"] = None if t_ivf is None else int(t_ivf * q_per_launch / q_ivf)
        line["roofline"]["traffic_source"] = tr.get("source")
        if lin is not None and "achieved" in lin:
            lin["traffic"] = tr.get("k_scan_stream32_linear_N64M_bytes_per_launch") if "N=64000000 " in lin["workload"] else None
    except Exception:
        pass
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------- reference ----
def reference_qps(cw, codes, Q, nproc, seconds_budget=20.0):
    """Reference (oracle/_ref/fast_*) throughput: index built by the reference itself, then `nproc` forked
    workers each looping single-query query_ivf calls over its slice of Q.  Returns (qps, n_queries, info)."""
    from oracle import ref as R
    kind = "fast"
    if not R.available(kind):
        return None
    r = R.Ref(kind)
    r.create(cw)
    r.add_codes(codes, False)
    t_rec = r.reconfigure(CFG["nlist"], CFG["iter"])
    # calibrate on a few queries, then size the sample for ~seconds_budget of wall time
    cal = r.time_queries(Q[:20], CFG["topk"], "ivf", L=CFG["L"])
    per_q = cal["seconds"] / cal["n"]
    n = int(min(len(Q), max(nproc * 8, seconds_budget / per_q * nproc)))
    res = r._call("time_queries_forked", Q=Q[:n], topk=CFG["topk"], L=CFG["L"], nproc=nproc)
    r.close()
    return res["n"] / res["seconds"], res["n"], {"kind": "reference", "build": r.kind, "reconfigure_s": round(t_rec, 2),
                                                 "single_thread_ms_per_query": round(per_q * 1e3, 4)}


def cpu_baseline_sample(cw, codes, Q):
    """cpu_baseline leg: the unmodified reference (oracle/_ref/fast_*) on the same codes and queries."""
    cores = os.cpu_count() or 1
    nproc = max(1, min(cores, 64))
    out = reference_qps(cw, codes, Q, nproc, seconds_budget=10.0)
    if out is None:
        return {"value": None, "unit": "queries/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    qps, n, info = out
    info.update({"value": round(qps, 1), "unit": "queries/s", "cores": nproc,
                 "sample": "%d single-query main.RiiCpp.query_ivf calls (L=32000, topk=1) over %d worker processes, same "
                           "codes/queries as the GPU arm (QueryIvf is single-threaded, src/rii.h:261,290)" % (n, nproc)})
    return info


def run_reference(args):
    """--impl reference: the unmodified reference (oracle/_ref/fast_*) on this box's host cores.  W warm-up steps and K timed
    steps; a step is a bounded sample of the C2 workload (same index, same L / topk): `sample` single-query calls spread over
    one worker process per core, sized from a calibration so that the whole run takes ~20-40 s."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import ref as R
    if not R.available("fast"):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/fast_* is not built on this box"}))
        return
    K, W = args.steps, args.warmup
    cw, codes, Q, _ = make_data_cpu(8192)
    cores = os.cpu_count() or 1
    nproc = max(1, min(cores, 64))
    r = R.Ref("fast")
    r.create(cw)
    r.add_codes(codes, False)
    t_rec = r.reconfigure(CFG["nlist"], CFG["iter"])
    cal = r.time_queries(Q[:20], CFG["topk"], "ivf", L=CFG["L"])
    per_q = cal["seconds"] / cal["n"]
    budget = 30.0
    sample = int(max(nproc * 4, min(len(Q), budget / (K + W) / per_q * nproc)))
    secs = []
    for i in range(W + K):
        o = (i * sample) % max(1, len(Q) - sample + 1)
        res = r._call("time_queries_forked", Q=Q[o:o + sample], topk=CFG["topk"], L=CFG["L"], nproc=nproc)
        if i >= W:
            secs.append(res["seconds"])
    r.close()
    total = sum(secs)
    qps = K * sample / total
    line = {"impl": "reference", "metric": "queries/sec at recall@1 (N=1M, D=128, M=32)", "value": round(qps, 1),
            "unit": "queries/s", "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": round(total / K * 1e3, 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic U[0,1)^128 float32 vectors, PQ trained on a 20k sample",
            "config": {"workload": "C2: N=1M D=128 M=32 Ks=256 IVF nlist=1000 L=32000 (w=35 lists) topk=1",
                       "batch_queries_per_step": sample,
                       "note": "a step = %d single-query calls (bounded sample of the same workload) over %d worker processes" % (sample, nproc)},
            "cpu_baseline": {"kind": "reference", "build": r.kind, "value": round(qps, 1), "unit": "queries/s", "cores": nproc,
                             "reconfigure_s": round(t_rec, 2), "single_thread_ms_per_query": round(per_q * 1e3, 4),
                             "sample": "%d steps x %d single-query main.RiiCpp.query_ivf calls over %d worker processes "
                                       "(QueryIvf is single-threaded, src/rii.h:261,290)" % (K, sample, nproc)},
            "e2e": {"value": round(qps, 1), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"  # before torch / NCCL load: rank 0 prints exactly one line, the JSON
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32768, help="queries per step")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--linear-n", type=int, default=1000000000,
                    help="also time the HBM-bound linear scan over this many random codes (0 = skip; falls back to 64M if "
                         "the GPU cannot hold it); 1 GPU only")
    ap.add_argument("--no-large", action="store_true", help="skip the C4 / C5 sharded legs (weak scaling, 12.5M / 125M codes per GPU)")
    ap.add_argument("--quick", action="store_true", help="skip the C3 subset leg")
    ap.add_argument("--gather", action="store_true", help="multi-GPU replicas: also all-gather the results so that every rank holds the whole batch "
                                                          "(one packed NCCL collective per step inside the timed region)")
    ap.add_argument("--shard", action="store_true", help="multi-GPU: partition the index by id range instead of replicating it")
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
