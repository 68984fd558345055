"""Id-range sharding of one Rii index over the GPUs of a box, one process per GPU.

Rank g owns the global ids [lo_g, hi_g) of `flattened_codes` and, of every posting list, the ids that fall in
its range (lists are ascending in id, so the concatenation over ranks is the list in stored order).
Replicated: codebooks, coarse centers, global list lengths.  Exchanges:
  build  : all-gather of the sampled codes (<= 100*nlist rows) and of the per-rank list lengths (nlist ints);
  search : all-gather of the per-shard top-k (k x (int64 id, float32 dist) per query) followed by a merge
           under (distance, id).  No other data-path collective exists: candidates are independent.

The orchestration is engine-agnostic (`ShardEngine` protocol) so that the host logic runs under gloo on CPU
in the tests with a stand-in engine; the product engine is `CudaShardEngine` (C ABI, NCCL).
"""
import ctypes as C

import numpy as np


def shard_bounds(n_total, world):
    return [(g * n_total) // world for g in range(world + 1)]


def reference_sample_ids(n_total, nlist):
    """First min(N, 100*nlist) ids of the reference's sampling shuffle (src/rii.h:115-124) via the C ABI."""
    from . import _capi
    n = C.c_int64(0)
    _capi.check(_capi.lib().rii_sample_ids(n_total, nlist, None, C.byref(n)))
    out = np.empty(n.value, np.int64)
    _capi.check(_capi.lib().rii_sample_ids(n_total, nlist, out.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(n)))
    return out


def gather_sample(local_codes, lo, hi, sample_ids, dist=None):
    """Assemble the PQk-means training sample (rows in shuffle order) from the shards that own the rows."""
    mask = (sample_ids >= lo) & (sample_ids < hi)
    pos = np.nonzero(mask)[0]
    part = (pos, np.ascontiguousarray(local_codes[sample_ids[mask] - lo]))
    if dist is None or dist.get_world_size() == 1:
        parts = [part]
    else:
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, part)
    M = local_codes.shape[1]
    sample = np.empty((len(sample_ids), M), np.uint8)
    for p, c in parts:
        sample[p] = c
    return sample


def exchange_lengths(local_len, rank, dist=None):
    """-> (global lengths, lengths held by lower ranks), both (nlist,) int32."""
    if dist is None or dist.get_world_size() == 1:
        return local_len.astype(np.int32), np.zeros_like(local_len, dtype=np.int32)
    alls = [None] * dist.get_world_size()
    dist.all_gather_object(alls, local_len.astype(np.int32))
    alls = np.stack(alls)
    return alls.sum(0).astype(np.int32), alls[:rank].sum(0).astype(np.int32)


def build_shard_generic(engine, local_codes, lo, n_total, nlist, iter, rank, dist, sample_ids_fn=reference_sample_ids):
    """The sharded equivalent of add_codes(all) + reconfigure(nlist, iter) (src/rii.h:108-156)."""
    hi = lo + local_codes.shape[0]
    engine.add_codes(local_codes)
    engine.set_shard(lo, n_total)
    ids = sample_ids_fn(n_total, nlist)
    sample = gather_sample(local_codes, lo, hi, ids, dist)
    centers = engine.fit_coarse(sample, nlist, iter)      # replicated, deterministic
    engine.set_coarse_centers(centers)                     # assigns the local codes
    glob, pre = exchange_lengths(engine.list_lengths(), rank, dist)
    engine.set_global_lengths(glob, pre)
    return centers


class CudaShardEngine(object):
    """`ShardEngine` over rii_b200.main.RiiCpp (one GPU)."""

    def __init__(self, impl):
        from . import _capi
        self.e, self.lib, self.check = impl, _capi.lib(), _capi.check

    def _p(self, a, t):
        return a.ctypes.data_as(C.POINTER(t))

    def add_codes(self, codes):
        self.e.add_codes(codes, False)

    def set_shard(self, lo, n_total):
        self.check(self.lib.rii_set_shard(self.e._h, int(lo), int(n_total)))

    def fit_coarse(self, sample, nlist, iter):
        sample = np.ascontiguousarray(sample, np.uint8)
        out = np.empty((nlist, self.e.M), np.uint8)
        self.check(self.lib.rii_fit_coarse(self.e._h, self._p(sample, C.c_uint8), sample.shape[0], nlist, iter,
                                           self._p(out, C.c_uint8)))
        return out

    def set_coarse_centers(self, centers):
        centers = np.ascontiguousarray(centers, np.uint8)
        self.check(self.lib.rii_set_coarse_centers(self.e._h, self._p(centers, C.c_uint8), centers.shape[0]))

    def list_lengths(self):
        out = np.empty(self.e.nlist, np.int32)
        self.check(self.lib.rii_copy_list_lengths(self.e._h, self._p(out, C.c_int32)))
        return out

    def set_global_lengths(self, glob, pre):
        glob, pre = np.ascontiguousarray(glob, np.int32), np.ascontiguousarray(pre, np.int32)
        self.check(self.lib.rii_set_global_lengths(self.e._h, self._p(glob, C.c_int32), self._p(pre, C.c_int32)))


def build_shard(impl, codes_all, nlist, iter, rank, world):
    """bench.py helper: every rank holds the whole (synthetic) code matrix and keeps its id range."""
    import torch.distributed as dist
    b = shard_bounds(codes_all.shape[0], world)
    return build_shard_generic(CudaShardEngine(impl), np.ascontiguousarray(codes_all[b[rank]:b[rank + 1]]), b[rank],
                               codes_all.shape[0], nlist, iter, rank, dist if world > 1 else None)


def sharded_query(engine, Q, topk, L, method, dist, world):
    """One batch on every shard, all-gather of the per-shard top-k, merge under (distance, id).
    `engine.query_local` returns torch tensors (ids int64 (B,k) global ids, dists float32 (B,k), counts int32 (B))
    on the engine's device; `engine.merge` takes the gathered (G,B,k)/(G,B) tensors."""
    ids, d, c = engine.query_local(Q, topk, L, method)
    return _gather_merge(engine, ids, d, c, dist, world)


def _gather_merge(engine, ids, d, c, dist, world):
    """all-gather of the per-shard top-k -- ONE collective: ids, distance bits and counts travel packed as int64 -- and the
    merge under (distance, id)."""
    import torch
    if world == 1:
        return ids, d, c
    B, k = ids.shape
    pack = torch.empty((B, 2 * k + 1), dtype=torch.int64, device=ids.device)
    pack[:, :k] = ids
    pack[:, k:2 * k] = d.view(torch.int32).to(torch.int64)
    pack[:, 2 * k] = c
    g = torch.empty((world, B, 2 * k + 1), dtype=torch.int64, device=ids.device)
    dist.all_gather_into_tensor(g.view(-1), pack.view(-1))
    g_ids = g[:, :, :k].contiguous()
    g_d = g[:, :, k:2 * k].to(torch.int32).view(torch.float32).contiguous()
    g_c = g[:, :, 2 * k].to(torch.int32).contiguous()
    return engine.merge(g_ids, g_d, g_c)


def sharded_query_split(engine, Q, topk, L, dist, world, rank):
    """Sharded IVF batch with the coarse phase split over the ranks (what large nlist needs: ranking 65536 centers costs
    as much as scanning a shard's candidates).  Rank r ranks the lists for queries [r*B/G, (r+1)*B/G), the (B, w) rankings
    are all-gathered, every rank scans its shard for all B queries, per-shard top-k are all-gathered and merged.  Queries
    whose plan is flagged (walk beyond w) are re-run through the unsplit path.  B must divide by world."""
    import torch
    B = Q.shape[0]
    assert B % world == 0
    Bl = B // world
    ranked_l = engine.coarse_rank(Q[rank * Bl:(rank + 1) * Bl], topk, L)           # (Bl, w) int32
    if world > 1:
        ranked = torch.empty((B, ranked_l.shape[1]), dtype=ranked_l.dtype, device=ranked_l.device)
        dist.all_gather_into_tensor(ranked.view(-1), ranked_l.contiguous().view(-1))
    else:
        ranked = ranked_l
    ids, d, c, flags = engine.query_ranked(Q, topk, L, ranked)
    redo = torch.nonzero(flags & 1).flatten()   # the plan is global: the same queries on every rank
    if redo.numel():
        i2, d2, c2 = engine.query_local(Q[redo].contiguous(), topk, L, "ivf")
        ids[redo], d[redo], c[redo] = i2, d2, c2
    return _gather_merge(engine, ids, d, c, dist, world)


def sharded_query_subset(engine, Q, topk, L, tids, dist, world, rank):
    """IVF + target_ids on id-range shards (the one exchange step; src/rii.h:286-322 with the
    binary_search filter of :294).  Every shard builds the sub-index of its own members of `tids` (the members of every
    posting list, ascending in id) and reports its per-list member counts; one all-gather of nlist counts; then the
    queries are ordinary sharded IVF searches over the sub-indexes -- the cut after L member candidates and the topk test
    at the w-th list are planned from the global counts, identically on every rank.
    `tids`: sorted global int64 ids (a tensor on the engine's device)."""
    import torch
    cnt = engine.subset_begin(tids)                                       # (nlist,) int32, this shard
    if world > 1:
        g = torch.empty((world,) + tuple(cnt.shape), dtype=cnt.dtype, device=cnt.device)
        dist.all_gather_into_tensor(g.view(-1), cnt.contiguous().view(-1))
        glob = g.sum(0, dtype=torch.int32).contiguous()
        pre = (g[:rank].sum(0, dtype=torch.int32) if rank > 0 else torch.zeros_like(cnt)).contiguous()
        engine.subset_set_global(glob, pre)
    ids, d, c = engine.subset_query(Q, topk, L)
    return _gather_merge(engine, ids, d, c, dist, world)


def _cuda_query_local(self, Q, topk, L, method):
    import torch
    B = Q.shape[0]
    dev = Q.device
    ids = torch.empty((B, topk), dtype=torch.int64, device=dev)
    d = torch.empty((B, topk), dtype=torch.float32, device=dev)
    c = torch.empty((B,), dtype=torch.int32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    self.check(self.lib.rii_query_batch_dev(self.e._h, C.c_void_p(Q.data_ptr()), B, topk, None, 0, int(L),
                                            {"linear": 0, "ivf": 1}[method], C.c_void_p(ids.data_ptr()),
                                            C.c_void_p(d.data_ptr()), C.c_void_p(c.data_ptr()), st))
    return ids, d, c


def _cuda_merge(self, g_ids, g_d, g_c):
    import torch
    G, B, k = g_ids.shape
    ids = torch.empty((B, k), dtype=torch.int64, device=g_ids.device)
    d = torch.empty((B, k), dtype=torch.float32, device=g_ids.device)
    c = torch.empty((B,), dtype=torch.int32, device=g_ids.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    self.check(self.lib.rii_merge_shards_dev(self.e._h, C.c_void_p(g_ids.data_ptr()), C.c_void_p(g_d.data_ptr()),
                                             C.c_void_p(g_c.data_ptr()), G, B, k, C.c_void_p(ids.data_ptr()),
                                             C.c_void_p(d.data_ptr()), C.c_void_p(c.data_ptr()), st))
    return ids, d, c


def _st():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _cuda_subset_begin(self, tids):
    import torch
    cnt = torch.empty((self.e.nlist,), dtype=torch.int32, device=tids.device)
    self.check(self.lib.rii_subset_begin_dev(self.e._h, C.c_void_p(tids.data_ptr()), tids.numel(), C.c_void_p(cnt.data_ptr()), _st()))
    return cnt


def _cuda_subset_set_global(self, glob, pre):
    self.check(self.lib.rii_subset_set_global_dev(self.e._h, C.c_void_p(glob.data_ptr()), C.c_void_p(pre.data_ptr()), _st()))


def _cuda_subset_query(self, Q, topk, L):
    import torch
    B, dev = Q.shape[0], Q.device
    ids = torch.empty((B, topk), dtype=torch.int64, device=dev)
    d = torch.empty((B, topk), dtype=torch.float32, device=dev)
    c = torch.empty((B,), dtype=torch.int32, device=dev)
    self.check(self.lib.rii_subset_query_dev(self.e._h, C.c_void_p(Q.data_ptr()), B, topk, int(L), C.c_void_p(ids.data_ptr()),
                                             C.c_void_p(d.data_ptr()), C.c_void_p(c.data_ptr()), _st()))
    return ids, d, c


def _cuda_coarse_rank(self, Q, topk, L):
    import torch
    w = self.check(self.lib.rii_coarse_width(self.e._h, int(L)))
    ranked = torch.empty((Q.shape[0], w), dtype=torch.int32, device=Q.device)
    if Q.shape[0]:
        self.check(self.lib.rii_coarse_rank_dev(self.e._h, C.c_void_p(Q.data_ptr()), Q.shape[0], topk, int(L),
                                                C.c_void_p(ranked.data_ptr()), _st()))
    return ranked


def _cuda_query_ranked(self, Q, topk, L, ranked):
    import torch
    B, dev = Q.shape[0], Q.device
    ids = torch.empty((B, topk), dtype=torch.int64, device=dev)
    d = torch.empty((B, topk), dtype=torch.float32, device=dev)
    c = torch.empty((B,), dtype=torch.int32, device=dev)
    flags = torch.empty((B,), dtype=torch.int32, device=dev)
    self.check(self.lib.rii_query_ranked_dev(self.e._h, C.c_void_p(Q.data_ptr()), B, topk, int(L), C.c_void_p(ranked.data_ptr()),
                                             C.c_void_p(ids.data_ptr()), C.c_void_p(d.data_ptr()), C.c_void_p(c.data_ptr()),
                                             C.c_void_p(flags.data_ptr()), _st()))
    return ids, d, c, flags


CudaShardEngine.query_local = _cuda_query_local
CudaShardEngine.subset_begin = _cuda_subset_begin
CudaShardEngine.subset_set_global = _cuda_subset_set_global
CudaShardEngine.subset_query = _cuda_subset_query
CudaShardEngine.coarse_rank = _cuda_coarse_rank
CudaShardEngine.query_ranked = _cuda_query_ranked
CudaShardEngine.merge = _cuda_merge


class LocalShardGroup(object):
    """G id-range shards held by ONE process as G index handles (all on one GPU, or spread over the GPUs the process
    sees).  Exactly the C-ABI sequence of the multi-process path -- rii_set_shard, rii_fit_coarse, rii_set_coarse_centers,
    rii_set_global_lengths, per-shard queries, rii_merge_shards_dev -- with every collective replaced by a concatenation
    on the host / device.  It is what the single-GPU parity test of the sharded planner drives (pre_len / glob_len), and
    a way to shard an index over several GPUs without torch.distributed."""

    def __init__(self, impls):
        self.engines = [CudaShardEngine(e) for e in impls]
        self.G = len(impls)

    def build(self, codes_all, nlist, iter):
        n_total = codes_all.shape[0]
        b = shard_bounds(n_total, self.G)
        ids = reference_sample_ids(n_total, nlist)
        for g, eng in enumerate(self.engines):
            eng.add_codes(np.ascontiguousarray(codes_all[b[g]:b[g + 1]]))
            eng.set_shard(b[g], n_total)
        sample = np.ascontiguousarray(codes_all[ids])        # == the concatenation gather_sample() assembles
        centers = self.engines[0].fit_coarse(sample, nlist, iter)
        lens = []
        for eng in self.engines:
            eng.set_coarse_centers(centers)
            lens.append(eng.list_lengths().astype(np.int64))
        lens = np.stack(lens)
        for g, eng in enumerate(self.engines):
            eng.set_global_lengths(lens.sum(0).astype(np.int32), lens[:g].sum(0).astype(np.int32))
        return centers

    def _merge(self, outs):
        import torch
        if self.G == 1:
            return outs[0]
        dev = outs[0][0].device
        g_ids = torch.stack([o[0].to(dev) for o in outs]).contiguous()
        g_d = torch.stack([o[1].to(dev) for o in outs]).contiguous()
        g_c = torch.stack([o[2].to(dev) for o in outs]).contiguous()
        return self.engines[0].merge(g_ids, g_d, g_c)

    def query(self, Q, topk, L, method):
        import torch
        outs = []
        for eng in self.engines:
            with torch.cuda.device(eng.e._device):
                outs.append(eng.query_local(Q.to("cuda:%d" % eng.e._device), topk, L, method))
                torch.cuda.synchronize()
        return self._merge(outs)

    def query_subset(self, Q, topk, L, tids):
        """IVF + target_ids over the shards: sub-index per shard, exchange of the per-list member counts, sharded IVF."""
        import torch
        cnts = torch.stack([eng.subset_begin(tids) for eng in self.engines])
        glob = cnts.sum(0, dtype=torch.int32).contiguous()
        outs = []
        for r, eng in enumerate(self.engines):
            pre = (cnts[:r].sum(0, dtype=torch.int32) if r else torch.zeros_like(glob)).contiguous()
            if self.G > 1:
                eng.subset_set_global(glob, pre)
            outs.append(eng.subset_query(Q, topk, L))
        torch.cuda.synchronize()
        return self._merge(outs)

    def query_split(self, Q, topk, L):
        """The coarse / scan split of sharded_query_split with the G handles of this process."""
        import torch
        B = Q.shape[0]
        assert B % self.G == 0
        Bl = B // self.G
        ranked = torch.cat([eng.coarse_rank(Q[r * Bl:(r + 1) * Bl].contiguous(), topk, L) for r, eng in enumerate(self.engines)])
        outs = []
        for eng in self.engines:
            ids, d, c, flags = eng.query_ranked(Q, topk, L, ranked)
            redo = torch.nonzero(flags & 1).flatten()
            if redo.numel():
                i2, d2, c2 = eng.query_local(Q[redo].contiguous(), topk, L, "ivf")
                ids[redo], d[redo], c[redo] = i2, d2, c2
            outs.append((ids, d, c))
        torch.cuda.synchronize()
        return self._merge(outs)
