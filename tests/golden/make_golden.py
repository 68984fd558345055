"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Runs only where /root/reference exists (it drives oracle/_ref/strict_*: the reference sources compiled as
written, -O2 -ffp-contract=off; see oracle/Makefile).  Usage:  python tests/golden/make_golden.py

The reference has no golden vectors of its own, so these fixtures *are* the pinned
behaviour of the path: for seeded inputs they record what `main.RiiCpp` returns --
  * every ADC distance (query_linear with topk = N), hence the distance table + ADist arithmetic,
  * query_linear / query_ivf results (raw, and canonicalised to the (distance, id) order by asking the
    reference for *all* candidates of the same candidate set),
  * reconfigure(): coarse centers and posting lists.
Each case is small (< 300 KB) and carries its inputs, so nothing depends on RNG reproducibility.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from oracle import ref as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name, D, M, Ks, N, nlist, iter, nq, builds
CASES = [
    ("c1_d128_m32", 128, 32, 256, 2000, 20, 5, 6, ["strict_v4"]),      # BASELINE C1/C2/C3 shape (Ds=4)
    ("t_d40_m4_ks20", 40, 4, 20, 1000, 20, 5, 6, ["strict_v4"]),       # the reference tests' shape (Ds=10)
    ("t_d40_m20", 40, 20, 256, 1000, 10, 3, 4, ["strict_v4"]),         # tests/test_rii.py:146 (M=20, Ds=2)
    ("c5_d96_m32", 96, 32, 256, 1500, 15, 2, 4, ["strict_v4"]),        # BASELINE C5 shape (Ds=3)
    ("c4_d128_m64", 128, 64, 256, 1500, 12, 2, 4, ["strict_v4"]),      # BASELINE C4 shape (Ds=2)
    ("w_d80_m4_ks16", 80, 4, 16, 600, 6, 2, 4, ["strict_v4", "strict_v3"]),  # Ds=20: 16- vs 8-lane fvec_L2sqr
]


def canonical(ids, dists, k):
    ids = np.asarray(ids, np.int64)
    dists = np.asarray(dists, np.float32)
    o = np.lexsort((ids, dists))[:k]
    return ids[o], dists[o]


def make_case(name, D, M, Ks, N, nlist, it, nq, build):
    import zlib
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    Ds = D // M
    # codewords and database vectors like the reference's tests/README (uniform [0,1) float32)
    cw = rng.random((M, Ks, Ds), dtype=np.float32)
    codes = rng.integers(0, Ks, (N, M), dtype=np.uint8)
    Q = rng.random((nq, D), dtype=np.float32)
    W = R.simd_width(build)
    r = R.Ref(build)
    r.create(cw)
    r.add_codes(codes, False)
    r.reconfigure(nlist, it)
    st = r.state()
    centers = st["coarse_centers"]
    offsets, ids = O.lists_to_csr(st["posting_lists"])
    out = dict(cw=cw, codes=codes, Q=Q, nlist=nlist, iter=it, variant=W, centers=centers, offsets=offsets, ids=ids)

    # every ADC distance
    all_d = np.empty((nq, N), np.float32)
    for i, q in enumerate(Q):
        rid, rd = r.query_linear(q, N, None)
        all_d[i, np.asarray(rid)] = np.asarray(rd, np.float32)
    out["all_dists"] = all_d

    # linear: raw top-k, subset
    tids = np.sort(rng.choice(N, N // 5, replace=False)).astype(np.int64)
    out["tids"] = tids
    lin = []
    for i, q in enumerate(Q):
        for topk in (1, 3, 50):
            rid, rd = r.query_linear(q, topk, None)
            sid, sd = r.query_linear(q, topk, tids)
            lin.append((i, topk, np.asarray(rid), np.asarray(rd, np.float32), np.asarray(sid), np.asarray(sd, np.float32)))
    out["lin_meta"] = np.array([(a, b) for a, b, *_ in lin], np.int64)
    for j, (_, _, rid, rd, sid, sd) in enumerate(lin):
        out["lin_%d_ids" % j], out["lin_%d_d" % j] = rid, rd
        out["lin_%d_sids" % j], out["lin_%d_sd" % j] = sid, sd

    # ivf: (topk, L, use_subset)
    L0 = int(np.round(N / nlist))
    combos = [(1, L0, 0), (3, 2 * L0, 0), (10, 5 * L0 + 7, 0), (5, N, 0), (1, L0, 1), (7, 3 * L0, 1), (4, len(tids), 1)]
    meta = []
    skipped = 0
    j = 0
    for i, q in enumerate(Q):
        T = O.dtable(q, cw, W)
        for topk, L, sub in combos:
            t = tids if sub else None
            if topk > L or L > N or (sub and topk > len(tids)):
                continue
            rid, rd = r.query_ivf(q, topk, t, L)
            oid, od, ncand = O.query_ivf(T, codes, centers, offsets, ids, topk, L, t, return_ncand=True)
            if len(rid) == 0:
                if len(oid) != 0:
                    skipped += 1
                    continue
                cid, cd = np.zeros(0, np.int64), np.zeros(0, np.float32)
            else:
                # all candidates of the same candidate set (see docstring) -> canonical (dist, id) top-k
                aid, ad = r.query_ivf(q, ncand, t, L) if ncand >= topk else ([], [])
                if len(aid) != ncand:
                    skipped += 1  # unspecified-order regime (walk beyond w) or a coarse tie: not pinned
                    continue
                cid, cd = canonical(aid, ad, topk)
            meta.append((i, topk, L, sub, ncand))
            out["ivf_%d_raw_ids" % j], out["ivf_%d_raw_d" % j] = np.asarray(rid, np.int64), np.asarray(rd, np.float32)
            out["ivf_%d_ids" % j], out["ivf_%d_d" % j] = cid, cd
            j += 1
    out["ivf_meta"] = np.array(meta, np.int64)
    r.close()
    path = os.path.join(HERE, "%s.%s.npz" % (name, build))
    np.savez_compressed(path, **out)
    print("%-28s %s  ivf cases %d (skipped %d)  %.0f KB" % (name, build, len(meta), skipped, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    if not os.path.isdir("/root/reference/src"):
        sys.exit("needs /root/reference (the fixtures are committed; regenerate only in the build container)")
    for name, D, M, Ks, N, nlist, it, nq, builds in CASES:
        for b in builds:
            if not R.available(b):
                sys.exit("oracle/_ref/%s missing: run `make -C oracle ref`" % b)
            make_case(name, D, M, Ks, N, nlist, it, nq, b)
