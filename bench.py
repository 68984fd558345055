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
             queries; the per-rank results are all-gathered over NCCL so that every rank holds the whole batch.
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
def make_data(nq, device):
    """Synthetic vectors, a PQ trained on a 20k sample (scipy k-means), codes by exact nearest codeword and the
    exact float-L2 ground truth for recall@1 -- all *setup* (torch on the GPU), outside any timed region."""
    import torch
    from rii_b200 import pq
    N, D, M, Ks = CFG["N"], CFG["D"], CFG["M"], CFG["Ks"]
    Ds = D // M
    g = torch.Generator(device=device).manual_seed(123)
    X = torch.rand((N, D), generator=g, device=device, dtype=torch.float32)
    g2 = torch.Generator(device=device).manual_seed(456)
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
        dist.init_process_group("nccl", device_id=dev)
    lib = _capi.lib()
    B, K, W = args.batch, args.steps, args.warmup
    nq = B * min(4, K + W)  # a few distinct query batches, cycled
    cw, codes, Q, gt = make_data(nq, dev)
    N, M, D = CFG["N"], CFG["M"], CFG["D"]

    # ---- index: id-range shard per rank -----------------------------------------------------------
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
    o_ids = torch.empty((B, k), dtype=torch.int64, device=dev)
    o_d = torch.empty((B, k), dtype=torch.float32, device=dev)
    o_c = torch.empty((B,), dtype=torch.int32, device=dev)
    Bl = B if (world == 1 or shard) else B // world  # queries this rank answers per step
    assert B % world == 0
    if world > 1:
        g_ids = torch.empty((world, Bl, k), dtype=torch.int64, device=dev)
        g_d = torch.empty((world, Bl, k), dtype=torch.float32, device=dev)
        g_c = torch.empty((world, Bl), dtype=torch.int32, device=dev)
        f_ids, f_d, f_c = torch.empty_like(o_ids), torch.empty_like(o_d), torch.empty_like(o_c)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_dev(i):
        q = dQ[(i % (nq // B)) * B:(i % (nq // B) + 1) * B]
        if world > 1 and not shard:  # replicas: this rank's slice of the batch, then all-gather of the results
            q = q[rank * Bl:(rank + 1) * Bl]
        _capi.check(lib.rii_query_batch_dev(e._h, C.c_void_p(q.data_ptr()), Bl, k, None, 0, L, 1,
                                            C.c_void_p(o_ids.data_ptr()), C.c_void_p(o_d.data_ptr()),
                                            C.c_void_p(o_c.data_ptr()), sp))
        if world > 1:
            dist.all_gather_into_tensor(g_ids.view(-1), o_ids[:Bl].reshape(-1))
            dist.all_gather_into_tensor(g_d.view(-1), o_d[:Bl].reshape(-1))
            dist.all_gather_into_tensor(g_c.view(-1), o_c[:Bl])
            if not shard:
                return g_ids.view(B, k)
            _capi.check(lib.rii_merge_shards_dev(e._h, C.c_void_p(g_ids.data_ptr()), C.c_void_p(g_d.data_ptr()),
                                                 C.c_void_p(g_c.data_ptr()), world, B, k,
                                                 C.c_void_p(f_ids.data_ptr()), C.c_void_p(f_d.data_ptr()),
                                                 C.c_void_p(f_c.data_ptr()), sp))
            return f_ids
        return o_ids

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
        ids = step_dev(W + i)
        b.record(st)
        evs.append((a, b))
        if i < nq // B:
            got.append((W + i, ids.clone()))
    barrier()
    launches = lib.rii_launch_count() - launches0
    sampler.stop_flag = True
    ms = sum(a.elapsed_time(b) for a, b in evs)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    scan_ms, scan_n = C.c_double(0), C.c_int64(0)
    lib.rii_profile_get(e._h, b"scan_ivf", C.byref(scan_ms), C.byref(scan_n))
    prof = {}
    for name in ("dtable", "coarse_rank", "scan_ivf", "merge"):
        m_, n_ = C.c_double(0), C.c_int64(0)
        lib.rii_profile_get(e._h, name.encode(), C.byref(m_), C.byref(n_))
        prof[name] = {"ms_total": round(m_.value, 4), "launches": n_.value}
    lib.rii_profile_enable(e._h, 0)
    sampler.join(timeout=2)

    # recall@1 (examples/benchmark/util.py:35-58) of what the timed steps returned
    hit = tot = 0
    for i, ids in got:
        s = (i % (nq // B)) * B
        hit += int((ids[:, 0].cpu().numpy() == gt[s:s + B]).sum())
        tot += B
    recall = hit / max(tot, 1)

    # ---- end-to-end arm: host (pinned) buffers through rii_query_batch ------------------------------
    e2e = None
    if world == 1:
        hQ = torch.from_numpy(Q).pin_memory()
        h_ids = torch.empty((B, k), dtype=torch.int64).pin_memory()
        h_d = torch.empty((B, k), dtype=torch.float32).pin_memory()
        h_c = torch.empty((B,), dtype=torch.int32).pin_memory()

        def step_host(i):
            q = hQ[(i % (nq // B)) * B:(i % (nq // B) + 1) * B]
            _capi.check(lib.rii_query_batch(e._h, C.cast(q.data_ptr(), C.POINTER(C.c_float)), B, k, None, 0, L, 1,
                                            C.cast(h_ids.data_ptr(), C.POINTER(C.c_int64)),
                                            C.cast(h_d.data_ptr(), C.POINTER(C.c_float)),
                                            C.cast(h_c.data_ptr(), C.POINTER(C.c_int32))))
        for i in range(W):
            step_host(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(K):
            step_host(W + i)  # synchronous: returns after the D2H of the results
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        # single-query latency through the reference's own call shape (query_ivf, one query per call)
        qs = Q[:200]
        for q in qs[:10]:
            e.query_ivf(q, k, np.empty(0, np.int64), L)
        t1 = time.perf_counter()
        for q in qs:
            e.query_ivf(q, k, np.empty(0, np.int64), L)
        lat = (time.perf_counter() - t1) / len(qs)
        e2e = {"value": round(K * B / dt, 1), "unit": "queries/s", "h2d_bytes_per_step": B * D * 4,
               "d2h_bytes_per_step": B * k * 12 + B * 4, "single_query_call_us": round(lat * 1e6, 1)}

    # ---- the HBM-bound side of the same engine: linear PQ-code scan over N >> L2 (north_star's ">= 70 % of the HBM
    # roofline on the code scan"); random codes (throughput only), 1 query per launch, CUDA events inside the library
    lin = None
    if world == 1 and args.linear_n > 0:
        try:
            e.set_option("fuse_coarse", 1)
            el = main.RiiCpp(cw, False, device=local, l2_variant=16)
            nl, chunk = int(args.linear_n), 8000000
            gen = torch.Generator(device=dev).manual_seed(7)
            for s0 in range(0, nl, chunk):
                c = min(chunk, nl - s0)
                el.add_codes(torch.randint(0, 256, (c, M), dtype=torch.uint8, device=dev, generator=gen).cpu().numpy(), False)
            lib.rii_profile_enable(el._h, 1)
            q1 = dQ[:1]
            li = torch.empty((1, 1), dtype=torch.int64, device=dev)
            ld = torch.empty((1, 1), dtype=torch.float32, device=dev)
            lc = torch.empty((1,), dtype=torch.int32, device=dev)
            for it in range(13):
                if it == 3:
                    torch.cuda.synchronize()
                    lib.rii_profile_reset(el._h)
                flush.zero_()
                _capi.check(lib.rii_query_batch_dev(el._h, C.c_void_p(q1.data_ptr()), 1, 1, None, 0, 0, 0,
                                                    C.c_void_p(li.data_ptr()), C.c_void_p(ld.data_ptr()),
                                                    C.c_void_p(lc.data_ptr()), sp))
            torch.cuda.synchronize()
            m_, n_ = C.c_double(0), C.c_int64(0)
            lib.rii_profile_get(el._h, b"scan_linear", C.byref(m_), C.byref(n_))
            lms = m_.value / max(n_.value, 1)
            lin = {"kernel": "k_scan_stream32<NW=12, linear, 4-stage rings, 1 CTA/SM>", "workload": "linear scan, N=%d M=32 (random codes), topk=1, 1 query/launch" % nl,
                   "bound": "hbm", "launch_ms": round(lms, 4), "algorithmic_bytes_per_launch": nl * M + 4 * M * CFG["Ks"],
                   "achieved": round((nl * M + 4 * M * CFG["Ks"]) / (lms * 1e-3) / 1e9, 1), "unit": "GB/s"}
            del el
        except Exception as ex:  # never lose the headline line over the side measurement
            lin = {"error": repr(ex)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    # algorithmic bytes of the dominant kernel (posting-list scan), per launch: per query C*M code bytes
    # (C = L candidates) + 4*M*Ks (its distance table).  SURVEY's V*4 bytes of visited ids are NOT counted: the
    # kernel streams a list-ordered (skew64) code copy and reads ids only for survivors; the nlist*M center bytes of
    # the fused coarse pass are not counted either (3 % of C*M at C2).
    frac_scanned = 1.0 / world if shard else 1.0
    q_per_launch = K * Bl / max(scan_n.value, 1)  # the library processes a step in chunks of <= 32768 queries
    alg = q_per_launch * (L * frac_scanned * M + 4 * M * CFG["Ks"])
    launch_ms = scan_ms.value / max(scan_n.value, 1)
    achieved = alg / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else None
    line = {
        "metric": "queries/sec at recall@1 (N=1M, D=128, M=32)", "value": round(K * B / (ms * 1e-3), 1),
        "unit": "queries/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(ms / K, 4),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic U[0,1)^128 float32 vectors, PQ trained on a 20k sample; ground truth exact L2",
        "config": {"workload": "C2: N=1M D=128 M=32 Ks=256 IVF nlist=1000 L=32000 (w=35 lists) topk=1",
                   "arithmetic": "uint8 codes index float32 tables; float32 adds in the reference's order (bit-exact)",
                   "batch_queries_per_step": B, "l2": "flushed between steps (256 MB write)",
                   "parallelism": "1 GPU" if world == 1 else
                   ("id-range shards x%d + NCCL all-gather of per-shard top-k + merge" % world if shard else
                    "index replicated x%d, queries split, NCCL all-gather of results" % world),
                   "index_build_s": round(t_build, 2)},
        "recall_at_1": round(recall, 4),
        "e2e": e2e, "gpu_launches": int(launches), "clocks": sampler.summary(),
        "roofline": {"kernel": "k_scan_stream32<NW=6, IVF fused (table + coarse + plan + scan), 3-stage rings, 2 CTAs/SM>", "bound": "hbm", "achieved": None if achieved is None else round(achieved, 1),
                     "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": None if achieved is None else round(achieved / peak, 4), "traffic": None,
                     "algorithmic_bytes_per_launch": int(alg), "queries_per_launch": int(q_per_launch),
                     "launch_ms": round(launch_ms, 4),
                     "note": "the 32 MB code table is L2-resident at N=1M: DRAM traffic << algorithmic bytes; the "
                             "binding resource is the shared-memory lookup rate (1 wavefront per 32 lookups) plus the per-query "
                             "serial phases; the HBM-bound case of the same engine is roofline_linear_scan (DESIGN.md)"},
        "kernel_ms": prof,
    }
    if lin is not None:
        if "achieved" in lin:
            lin.update({"peak": peak, "frac": round(lin["achieved"] / peak, 4)})
        line["roofline_linear_scan"] = lin
    if args.cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_sample(cw, codes, Q)
    try:  # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture (tools/ncu_summary.py)
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        # (captured on a launch of 8192 queries; DRAM traffic of this L2-resident workload is per query: scale to the
        # queries one launch of this run processes)
        t_ivf, q_ivf = tr.get("k_scan_stream32_ivf_2cta_bytes_per_launch"), tr.get("k_scan_stream32_ivf_2cta_queries_per_launch", 8192)
        line["roofline"]["traffic"] = None if t_ivf is None else int(t_ivf * q_per_launch / q_ivf)
        line["roofline"]["traffic_source"] = tr.get("source")
        if lin is not None and "achieved" in lin:
            lin["traffic"] = tr.get("k_scan_stream32_linear_N64M_bytes_per_launch") if int(args.linear_n) == 64000000 else None
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
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import ref as R
    if not R.available("fast"):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/fast_* is not built on this box"}))
        return
    B, K, W = args.batch, args.steps, args.warmup
    cw, codes, Q, _ = make_data_cpu(4096)
    cores = os.cpu_count() or 1
    nproc = max(1, min(cores, 64))
    qps, n, info = reference_qps(cw, codes, Q, nproc, seconds_budget=max(10.0, 4.0 * K))
    line = {"impl": "reference", "metric": "queries/sec at recall@1 (N=1M, D=128, M=32)", "value": round(qps, 1),
            "unit": "queries/s", "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": round(B / qps * 1e3, 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic U[0,1)^128 float32 vectors, PQ trained on a 20k sample",
            "config": {"workload": "C2: N=1M D=128 M=32 Ks=256 IVF nlist=1000 L=32000 (w=35 lists) topk=1",
                       "batch_queries_per_step": B},
            "cpu_baseline": dict(info, value=round(qps, 1), unit="queries/s", cores=nproc,
                                 sample="%d single-query main.RiiCpp.query_ivf calls over %d worker processes "
                                        "(QueryIvf is single-threaded, src/rii.h:261,290)" % (n, nproc)),
            "e2e": {"value": round(qps, 1), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32768, help="queries per step")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--linear-n", type=int, default=64000000,
                    help="also time the HBM-bound linear scan over this many random codes (0 = skip); 1 GPU only")
    ap.add_argument("--shard", action="store_true", help="multi-GPU: partition the index by id range instead of replicating it")
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
