"""world_size-2 gloo worker for tests/test_sharded_gloo.py: runs rii_b200.sharded's orchestration (build +
search) with an oracle-backed CPU engine and checks the merged results against the unsharded oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import O, synth  # noqa: E402
from rii_b200 import sharded  # noqa: E402


class OracleShardEngine(object):
    """CPU stand-in with the ShardEngine interface; the IVF plan is restated here in numpy (prefix-sum form,
    SURVEY A.3) independently of the CUDA make_plan()."""

    def __init__(self, cw):
        self.cw, self.M = cw, cw.shape[0]
        self.Dm = O.sym_matrices(cw)

    def add_codes(self, codes):
        self.codes = codes

    def set_shard(self, lo, n_total):
        self.lo, self.n_total = lo, n_total

    def set_coarse_centers(self, centers):
        self.centers = centers
        a = O.assign(self.Dm, self.codes, centers)
        self.offsets, self.ids = O.assign_to_lists(a, centers.shape[0])

    def list_lengths(self):
        return np.diff(self.offsets).astype(np.int32)

    def set_global_lengths(self, glob, pre):
        self.glob, self.pre = glob, pre

    def query_local(self, Q, topk, L, method):
        Q = Q.numpy()
        B = Q.shape[0]
        ids = np.full((B, topk), -1, np.int64)
        d = np.full((B, topk), np.inf, np.float32)
        c = np.zeros(B, np.int32)
        for b, q in enumerate(Q):
            T = O.dtable(q, self.cw, 16)
            if method == "linear":
                i, dd = O.query_linear(T, self.codes, min(topk, len(self.codes)))
            else:
                cand = self._ivf_candidates(T, topk, L)
                if cand is None:
                    continue
                dd_all = O.adist_all(T, self.codes[cand]) if len(cand) else np.zeros(0, np.float32)
                o = np.lexsort((cand, dd_all))[:topk]
                i, dd = cand[o], dd_all[o]
            n = len(i)
            ids[b, :n], d[b, :n], c[b] = np.asarray(i) + self.lo, dd, n
        return torch.from_numpy(ids), torch.from_numpy(d), torch.from_numpy(c)

    def _ivf_candidates(self, T, topk, L):
        nlist = self.centers.shape[0]
        cd = O.adist_all(T, self.centers)
        order = np.lexsort((np.arange(nlist), cd))
        w = min(int(np.floor(L * nlist / self.n_total + 0.5)) + 3, nlist)
        P, cand, done = 0, [], False
        for j, no in enumerate(order):
            f = int(self.glob[no])
            take = f
            if P + f >= L:
                take, done = L - P, True
            P += take
            lt = int(np.clip(take - self.pre[no], 0, self.offsets[no + 1] - self.offsets[no]))
            cand.append(self.ids[self.offsets[no]:self.offsets[no] + lt])
            if done or (j == w - 1 and P >= topk):
                done = True
                break
        if not done:
            return None
        return np.concatenate(cand).astype(np.int64) if cand else np.zeros(0, np.int64)

    # ---- IVF + target_ids: sub-index of the members + exchange of per-list member counts (restated in numpy
    # independently of the CUDA pipeline) ----
    def subset_begin(self, tids):
        t = np.unique(tids.numpy()) - self.lo
        t = t[(t >= 0) & (t < len(self.codes))]
        member = np.zeros(len(self.codes), bool)
        member[t] = True
        self._S = len(tids)
        self._sub = [self.ids[self.offsets[no]:self.offsets[no + 1]] for no in range(self.centers.shape[0])]
        self._sub = [m[member[m]] for m in self._sub]
        cnt = np.array([len(m) for m in self._sub], np.int32)
        self._sglob, self._spre = cnt.copy(), np.zeros_like(cnt)
        return torch.from_numpy(cnt)

    def subset_set_global(self, glob, pre):
        self._sglob, self._spre = glob.numpy(), pre.numpy()

    def subset_query(self, Q, topk, L):
        B = Q.shape[0]
        ids = np.full((B, topk), -1, np.int64)
        d = np.full((B, topk), np.inf, np.float32)
        c = np.zeros(B, np.int32)
        nlist = self.centers.shape[0]
        for b, q in enumerate(Q.numpy()):
            T = O.dtable(q, self.cw, 16)
            cd = O.adist_all(T, self.centers)
            order = np.lexsort((np.arange(nlist), cd))
            w = min(int(np.floor(L * nlist / self._S + 0.5)) + 3, nlist)
            # the reference's sequential walk (src/rii.h:286-322): stop at L, or after the w-th list with >= topk found
            P, cand, done = 0, [], False
            for j, no in enumerate(order):
                f = int(self._sglob[no])
                take = f
                if P + f >= L:
                    take, done = L - P, True
                P += take
                lt = int(np.clip(take - int(self._spre[no]), 0, len(self._sub[no])))
                cand.append(self._sub[no][:lt])
                if done or (j == w - 1 and P >= topk):
                    done = True
                    break
            if not done:
                continue  # src/rii.h:325: empty result
            cand = np.concatenate(cand).astype(np.int64) if cand else np.zeros(0, np.int64)
            dd = O.adist_all(T, self.codes[cand]) if len(cand) else np.zeros(0, np.float32)
            o = np.lexsort((cand, dd))[:topk]
            ids[b, :len(o)], d[b, :len(o)], c[b] = cand[o] + self.lo, dd[o], len(o)
        return torch.from_numpy(ids), torch.from_numpy(d), torch.from_numpy(c)

    # ---- coarse phase split from the scan ----
    def coarse_rank(self, Q, topk, L):
        nlist = self.centers.shape[0]
        w = min(int(np.floor(L * nlist / self.n_total + 0.5)) + 3, nlist)
        out = np.zeros((Q.shape[0], w), np.int32)
        for b, q in enumerate(Q.numpy()):
            cd = O.adist_all(O.dtable(q, self.cw, 16), self.centers)
            out[b] = np.lexsort((np.arange(nlist), cd))[:w]
        return torch.from_numpy(out)

    def query_ranked(self, Q, topk, L, ranked):
        B = Q.shape[0]
        ids = np.full((B, topk), -1, np.int64)
        d = np.full((B, topk), np.inf, np.float32)
        c = np.zeros(B, np.int32)
        flags = np.zeros(B, np.int32)
        nlist = self.centers.shape[0]
        w = ranked.shape[1]
        for b, q in enumerate(Q.numpy()):
            T = O.dtable(q, self.cw, 16)
            P, cand, done = 0, [], False
            for j, no in enumerate(ranked[b].numpy()):
                f = int(self.glob[no])
                take = f
                if P + f >= L:
                    take, done = L - P, True
                P += take
                lt = int(np.clip(take - self.pre[no], 0, self.offsets[no + 1] - self.offsets[no]))
                cand.append(self.ids[self.offsets[no]:self.offsets[no] + lt])
                if done or (j == w - 1 and P >= topk):
                    done = True
                    break
            if not done:
                flags[b] = 2 if w >= nlist else 1
                continue
            cand = np.concatenate(cand).astype(np.int64) if cand else np.zeros(0, np.int64)
            dd = O.adist_all(T, self.codes[cand]) if len(cand) else np.zeros(0, np.float32)
            o = np.lexsort((cand, dd))[:topk]
            ids[b, :len(o)], d[b, :len(o)], c[b] = cand[o] + self.lo, dd[o], len(o)
        return torch.from_numpy(ids), torch.from_numpy(d), torch.from_numpy(c), torch.from_numpy(flags)

    def merge(self, g_ids, g_d, g_c):
        G, B, k = g_ids.shape
        ids = torch.full((B, k), -1, dtype=torch.int64)
        d = torch.full((B, k), float("inf"))
        c = torch.zeros(B, dtype=torch.int32)
        for b in range(B):
            ii = np.concatenate([g_ids[g, b, :g_c[g, b]].numpy() for g in range(G)])
            dd = np.concatenate([g_d[g, b, :g_c[g, b]].numpy() for g in range(G)])
            o = np.lexsort((ii, dd))[:k]
            ids[b, :len(o)], d[b, :len(o)], c[b] = torch.from_numpy(ii[o]), torch.from_numpy(dd[o]), len(o)
        return ids, d, c


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D, M, Ks, N, nlist = 64, 8, 32, 6000, 24
    cw, codes, Q = synth(D, M, Ks, N, 6, seed=42)
    # unsharded expectation from the oracle (same on every rank)
    centers, assign = O.reconfigure(cw, codes, nlist, 3)
    offsets, ids = O.assign_to_lists(assign, nlist)
    b = sharded.shard_bounds(N, world)
    eng = OracleShardEngine(cw)
    eng.fit_coarse = lambda sample, nl, it: centers  # PQk-means itself is covered by the golden tests; here: plumbing
    got_centers = sharded.build_shard_generic(eng, codes[b[rank]:b[rank + 1]], b[rank], N, nlist, 3, rank, dist)
    assert np.array_equal(got_centers, centers)
    # the assembled training sample must be the reference's sample (src/rii.h:115-132) on every rank
    sid = sharded.reference_sample_ids(N, nlist)
    sample = sharded.gather_sample(codes[b[rank]:b[rank + 1]], b[rank], b[rank + 1], sid, dist)
    assert np.array_equal(sample, codes[sid])
    # lists: local ids ascending, global lengths consistent with the unsharded index
    assert np.array_equal(eng.glob, np.diff(offsets))
    alls = [None] * world
    dist.all_gather_object(alls, (eng.offsets, eng.ids))
    for no in range(nlist):
        cat = np.concatenate([i[o[no]:o[no + 1]] + b[g] for g, (o, i) in enumerate(alls)])
        assert np.array_equal(cat, ids[offsets[no]:offsets[no + 1]]), "list %d" % no
    # search
    Qt = torch.from_numpy(Q)
    for method, topk, L in [("linear", 1, 0), ("linear", 25, 0), ("ivf", 1, 250), ("ivf", 10, 1500), ("ivf", 40, 45),
                            ("ivf", 7, N)]:
        gi, gd, gc = sharded.sharded_query(eng, Qt, topk, L, method, dist, world)
        for bq, q in enumerate(Q):
            T = O.dtable(q, cw, 16)
            exp = O.query_linear(T, codes, topk) if method == "linear" else O.query_ivf(T, codes, centers, offsets, ids, topk, L)
            n = int(gc[bq])
            assert n == len(exp[0]), (method, topk, L, n, len(exp[0]))
            assert np.array_equal(gi[bq, :n].numpy(), exp[0]) and np.array_equal(gd[bq, :n].numpy().view(np.uint32), exp[1].view(np.uint32)), (method, topk, L)
    # the coarse phase split over the ranks (each rank ranks the lists for its half of the queries)
    for topk, L in [(1, 250), (10, 1500), (40, 45), (7, N)]:
        gi, gd, gc = sharded.sharded_query_split(eng, Qt, topk, L, dist, world, rank)
        for bq, q in enumerate(Q):
            exp = O.query_ivf(O.dtable(q, cw, 16), codes, centers, offsets, ids, topk, L)
            n = int(gc[bq])
            assert n == len(exp[0]), ("split", topk, L, n, len(exp[0]))
            assert np.array_equal(gi[bq, :n].numpy(), exp[0]) and np.array_equal(gd[bq, :n].numpy().view(np.uint32), exp[1].view(np.uint32)), ("split", topk, L)
    # IVF + target_ids across shards: uniform subset, a subset concentrated in few lists far from the queries (walk
    # beyond w -> flagged re-run with the full ranking), and one where L is never reached (empty result)
    rng = np.random.default_rng(5)
    far = np.argsort(O.adist_all(O.dtable(Q[0], cw, 16), centers))[-4:]
    far_ids = np.sort(np.concatenate([ids[offsets[no]:offsets[no + 1]] for no in far])).astype(np.int64)
    for tids, topk, L in [(np.sort(rng.choice(N, 900, replace=False)).astype(np.int64), 5, 300),
                          (np.sort(rng.choice(N, 900, replace=False)).astype(np.int64), 1, 900),
                          (far_ids, 3, 20), (far_ids[:6], 5, 6)]:
        gi, gd, gc = sharded.sharded_query_subset(eng, Qt, topk, L, torch.from_numpy(tids), dist, world, rank)
        for bq, q in enumerate(Q):
            exp = O.query_ivf(O.dtable(q, cw, 16), codes, centers, offsets, ids, topk, L, tids)
            n = int(gc[bq])
            assert n == len(exp[0]), ("subset", topk, L, n, len(exp[0]))
            assert np.array_equal(gi[bq, :n].numpy(), exp[0]) and np.array_equal(gd[bq, :n].numpy().view(np.uint32), exp[1].view(np.uint32)), ("subset", topk, L)
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok" % rank)


if __name__ == "__main__":
    main()
