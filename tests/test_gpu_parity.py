"""GPU parity tests proper: the CUDA path (through the C ABI, rii_b200.main.RiiCpp) against
  (1) the golden vectors recorded from the unmodified reference (tests/golden/),
  (2) the oracle restatement on fresh seeded inputs (sizes the oracle finishes in seconds),
  (3) size-independent properties at BASELINE.json's N = 1M.
Bar: ids bit-exact under the (distance, id) order, distances bit-exact fp32 (tolerance 0; the contract in
BASELINE.json is 1e-5 relative)."""
import numpy as np
import pytest

from _util import O, assert_same_result, bits, canonical_topk, golden_files, load_golden, synth

pytestmark = pytest.mark.gpu

FILES = golden_files("strict_v4") + golden_files("strict_v3")
EMPTY = np.array([], np.int64)


def engine(cw, codes=None, variant=16):
    from rii_b200 import main
    e = main.RiiCpp(cw, False, l2_variant=variant)
    if codes is not None:
        e.add_codes(codes, False)
    return e


def load_state(e, g):
    import ctypes as C
    from rii_b200 import _capi
    centers, codes = np.ascontiguousarray(g["centers"]), np.ascontiguousarray(g["codes"])
    offsets, ids = np.ascontiguousarray(g["offsets"], np.int64), np.ascontiguousarray(g["ids"], np.int32)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    _capi.check(_capi.lib().rii_set_state(e._h, p(centers, C.c_uint8), centers.shape[0], p(codes, C.c_uint8),
                                          codes.shape[0], p(offsets, C.c_int64), p(ids, C.c_int32)))


# ---------------------------------------------------------------- golden vectors (reference-derived) --
@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1])
def test_golden_dtable_adist(path):
    g = load_golden(path)
    e = engine(g["cw"], g["codes"], g["variant"])
    T = e.dtable(g["Q"])
    for i, q in enumerate(g["Q"]):
        assert np.array_equal(bits(T[i]), bits(O.dtable(q, g["cw"], g["variant"]))), "dtable %d" % i
        assert np.array_equal(bits(e.adist_all(q)), bits(g["all_dists"][i])), "adist %d" % i


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1])
def test_golden_query_linear(path):
    g = load_golden(path)
    e = engine(g["cw"], g["codes"], g["variant"])
    N = len(g["codes"])
    for j, (i, topk) in enumerate(g["lin_meta"]):
        q = g["Q"][i]
        ids, d = e.query_linear(q, int(topk), EMPTY)
        eid, ed = canonical_topk(np.arange(N), g["all_dists"][i], int(topk))
        assert_same_result(ids, np.array(d, np.float32), eid, ed, "linear")
        assert np.array_equal(bits(np.sort(g["lin_%d_d" % j])), bits(np.array(d, np.float32)))
        sids, sd = e.query_linear(q, int(topk), g["tids"])
        eid, ed = canonical_topk(g["tids"], g["all_dists"][i][g["tids"]], int(topk))
        assert_same_result(sids, np.array(sd, np.float32), eid, ed, "linear subset")


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1])
def test_golden_query_ivf(path):
    g = load_golden(path)
    e = engine(g["cw"], variant=g["variant"])
    load_state(e, g)
    for j, (i, topk, L, sub, ncand) in enumerate(g["ivf_meta"]):
        ids, d = e.query_ivf(g["Q"][i], int(topk), g["tids"] if sub else EMPTY, int(L))
        assert_same_result(ids, np.array(d, np.float32), g["ivf_%d_ids" % j], g["ivf_%d_d" % j],
                           "ivf case %d (topk=%d L=%d sub=%d)" % (j, topk, L, sub))


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1])
def test_golden_reconfigure(path):
    g = load_golden(path)
    e = engine(g["cw"], g["codes"], g["variant"])
    e.reconfigure(g["nlist"], g["iter"])
    assert np.array_equal(e.coarse_centers_array(), g["centers"])
    offsets, ids = e.posting_lists_csr()
    assert np.array_equal(offsets, g["offsets"])
    assert np.array_equal(ids, g["ids"])


# ---------------------------------------------------------------- oracle on fresh inputs ---------------
SHAPES = [(128, 32, 256), (40, 4, 20), (40, 20, 256), (96, 32, 256), (128, 64, 256), (64, 8, 64), (48, 16, 256),
          (36, 3, 7), (72, 4, 16)]


@pytest.mark.parametrize("D,M,Ks", SHAPES)
def test_oracle_linear_and_subset(D, M, Ks):
    N = 20000
    cw, codes, Q = synth(D, M, Ks, N, 4, seed=D * 1000 + M)
    e = engine(cw, codes)
    rng = np.random.default_rng(7)
    tids = np.sort(rng.choice(N, 3000, replace=False)).astype(np.int64)
    unsorted = rng.permutation(tids)
    for q in Q:
        T = O.dtable(q, cw, 16)
        for topk in (1, 10, 100, 1500):
            ids, d = e.query_linear(q, topk, EMPTY)
            assert_same_result(ids, np.array(d, np.float32), *O.query_linear(T, codes, topk), "linear k=%d" % topk)
            ids, d = e.query_linear(q, topk, tids)
            assert_same_result(ids, np.array(d, np.float32), *O.query_linear(T, codes, topk, tids), "subset")
        ids, d = e.query_linear(q, 10, unsorted)  # target ids need not be sorted for the linear scan
        assert_same_result(ids, np.array(d, np.float32), *O.query_linear(T, codes, 10, unsorted), "unsorted subset")


@pytest.mark.parametrize("D,M,Ks", [(128, 32, 256), (40, 4, 20), (40, 20, 256), (128, 64, 256)])
def test_oracle_reconfigure_assign_ivf(D, M, Ks):
    N, nlist = 30000, 60
    cw, codes, Q = synth(D, M, Ks, N, 6, seed=D + M)
    e = engine(cw, codes)
    e.reconfigure(nlist, 3)
    centers, assign = O.reconfigure(cw, codes, nlist, 3)
    assert np.array_equal(e.coarse_centers_array(), centers)
    offsets, ids = O.assign_to_lists(assign, nlist)
    eo, ei = e.posting_lists_csr()
    assert np.array_equal(eo, offsets) and np.array_equal(ei, ids)
    # K6 building block, with distances
    Dm = O.sym_matrices(cw)
    assert np.array_equal(bits(e.sym_matrices()), bits(Dm))
    a, d = e.assign(codes[:5000], centers, return_dist=True)
    oa, od = O.assign(Dm, codes[:5000], centers, return_dist=True)
    assert np.array_equal(a, oa) and np.array_equal(bits(d), bits(od))
    # IVF
    rng = np.random.default_rng(3)
    tids = np.sort(rng.choice(N, N // 10, replace=False)).astype(np.int64)
    L0 = N // nlist
    for q in Q:
        T = O.dtable(q, cw, 16)
        for topk, L, t in [(1, L0, None), (10, 4 * L0, None), (100, 32 * L0 // 4, None), (5, N, None), (1, L0, tids),
                           (20, 3 * L0, tids), (3, len(tids), tids), (50, 17, None) if False else (17, 50, None)]:
            ids, dd = e.query_ivf(q, topk, EMPTY if t is None else t, L)
            oi, odd = O.query_ivf(T, codes, centers, offsets, ids=ei, topk=topk, L=L, tids=t)
            assert_same_result(ids, np.array(dd, np.float32), oi, odd, "ivf k=%d L=%d sub=%s" % (topk, L, t is not None))


def test_ivf_walk_beyond_w_and_empty():
    """SURVEY A.3 3rd/4th bullet: target ids concentrated in far lists -> the walk continues beyond w (we
    re-run with the full ranking, oracle order (dist, list id)), or nothing is found at all (empty result)."""
    D, M, Ks, N, nlist = 128, 32, 256, 20000, 40
    cw, codes, Q = synth(D, M, Ks, N, 4, seed=99)
    e = engine(cw, codes)
    e.reconfigure(nlist, 2)
    centers = e.coarse_centers_array()
    offsets, ids = e.posting_lists_csr()
    for q in Q:
        T = O.dtable(q, cw, 16)
        cd = np.array([O.adist_all(T, centers)]).reshape(-1)
        far = np.argsort(cd)[-6:]
        tids = np.sort(np.concatenate([ids[offsets[n]:offsets[n + 1]] for n in far])).astype(np.int64)
        for topk, L in [(5, 20), (3, len(tids)), (len(tids), len(tids))]:
            got = e.query_ivf(q, topk, tids, L)
            exp = O.query_ivf(T, codes, centers, offsets, ids, topk, L, tids)
            assert_same_result(got[0], np.array(got[1], np.float32), exp[0], exp[1], "beyond-w k=%d L=%d" % (topk, L))
        # L can never be reached and fewer than topk ids are in the first w lists -> empty (src/rii.h:325)
        few = tids[:3]
        got = e.query_ivf(q, 3, few, 3)
        exp = O.query_ivf(T, codes, centers, offsets, ids, 3, 3, few)
        assert_same_result(got[0], np.array(got[1], np.float32), exp[0], exp[1], "tiny subset")


def test_add_update_merge_clear_pickle():
    import copy
    import pickle
    D, M, Ks = 128, 32, 256
    cw, codes, Q = synth(D, M, Ks, 12000, 3, seed=5)
    e = engine(cw, codes[:8000])
    e.reconfigure(30, 2)
    e.add_codes(codes[8000:], True)  # src/rii.h:189-192
    assert e.N == 12000
    centers = e.coarse_centers_array()
    Dm = O.sym_matrices(cw)
    assign = np.concatenate([O.assign(Dm, codes[:8000], centers), O.assign(Dm, codes[8000:], centers)])
    offsets, ids = O.assign_to_lists(assign, 30)
    eo, ei = e.posting_lists_csr()
    assert np.array_equal(eo, offsets) and np.array_equal(ei, ids)
    assert np.array_equal(e.codes_array(), codes)
    e2 = pickle.loads(pickle.dumps(e))
    e3 = copy.deepcopy(e)
    for x in (e2, e3):
        assert x.N == e.N and x.nlist == e.nlist
        assert x.posting_lists == e.posting_lists and x.coarse_centers == e.coarse_centers
        assert x.query_ivf(Q[0], 5, EMPTY, 800) == e.query_ivf(Q[0], 5, EMPTY, 800)
    e.clear()
    assert e.N == 0 and e.nlist == 0 and e.posting_lists == [] and e.coarse_centers == []
    with pytest.raises(Exception):
        e.add_codes(codes[:10], True)  # no coarse centers: the reference terminates (src/rii.h:166-170)


def test_edge_cases():
    D, M, Ks = 128, 32, 256
    cw, codes, Q = synth(D, M, Ks, 1000, 2, seed=11)
    e = engine(cw, codes[:1])
    assert e.query_linear(Q[0], 1, EMPTY)[0] == [0]
    e.reconfigure(1, 5)
    assert e.posting_lists == [[0]]
    assert e.query_ivf(Q[0], 1, EMPTY, 1)[0] == [0]
    e = engine(cw, codes)
    T = O.dtable(Q[0], cw, 16)
    ids, d = e.query_linear(Q[0], 1000, EMPTY)  # topk == N (maximum)
    assert_same_result(ids, np.array(d, np.float32), *O.query_linear(T, codes, 1000), "topk=N")
    dup = np.array([5, 5, 9, 5], np.int64)  # duplicates -> duplicate results (docs: undefined; we keep them)
    ids, d = e.query_linear(Q[0], 4, dup)
    assert sorted(ids) == [5, 5, 5, 9]
    with pytest.raises(TypeError):
        e.query_linear(Q[0].astype(np.float64), 1, EMPTY)  # .noconvert(), src/main.cpp:18
    with pytest.raises(TypeError):
        e.query_linear(Q[0], 1, np.array([1], np.int32))
    with pytest.raises(ValueError):
        e.query_linear(Q[0], 1001, EMPTY)  # topk > N: the reference asserts (src/rii.h:200)
    with pytest.raises(ValueError):
        e.query_linear(Q[0], 1, np.array([1000], np.int64))  # out-of-range id: UB in the reference, refused here
    with pytest.raises(Exception):
        e.query_ivf(Q[0], 1, EMPTY, 10)  # no posting lists yet
    ids, d, counts = e.query_batch(np.ascontiguousarray(Q), 7, method="linear")
    for b in range(2):
        exp = O.query_linear(O.dtable(Q[b], cw, 16), codes, 7)
        assert_same_result(ids[b], d[b], exp[0], exp[1], "batch")
    assert counts.tolist() == [7, 7]


# ---------------------------------------------------------------- the reference's own property tests --
def test_reference_properties_public_api():
    """tests/test_rii.py:117-218 (test_query_linear / test_query_ivf / test_query) against rii_b200.Rii."""
    import rii_b200 as rii
    np.random.seed(123)
    M, Ks, N, D = 20, 256, 1000, 40
    X = np.random.random((N, D)).astype(np.float32)
    e = rii.Rii(fine_quantizer=rii.PQ(M=M, Ks=Ks, verbose=False).fit(vecs=X, iter=5)).add_configure(vecs=X, nlist=20)
    all_ids = np.arange(N, dtype=np.int64)
    S = np.sort(np.random.choice(N, 300, replace=False)).astype(np.int64)
    for n, q in enumerate(X[:20]):
        ids1, d1 = e.impl_cpp.query_linear(q, 10, EMPTY)
        assert isinstance(ids1, list) and isinstance(ids1[0], int) and isinstance(d1[0], float) and len(ids1) == 10
        assert all(d1[i] <= d1[i + 1] for i in range(9)) and n in ids1
        assert (ids1, d1) == tuple(e.impl_cpp.query_linear(q, 10, all_ids))        # :138-140
        ids_s, _ = e.impl_cpp.query_linear(q, 10, S)
        assert set(ids_s) <= set(S.tolist())                                       # :143-144
        assert tuple(e.impl_cpp.query_ivf(q, 10, all_ids, N)) == (ids1, d1)        # :178-181
        assert tuple(e.impl_cpp.query_ivf(q, 10, S, len(S))) == tuple(e.impl_cpp.query_linear(q, 10, S))  # :183-187
        ids, dists = e.query(q, topk=50, method="ivf", L=600)
        assert ids.dtype == np.int64 and dists.dtype == np.float64 and len(ids) == 50
        assert np.all(np.diff(dists) >= 0)
        ids, dists = e.query(q, topk=3)  # method='auto' goes through the fitted threshold
        assert len(ids) == 3
    # OPQ path: query is rotated first (rii/rii.py:305-306)
    eo = rii.Rii(fine_quantizer=rii.OPQ(M=4, Ks=20, verbose=False).fit(vecs=X, pq_iter=5, rotation_iter=3))
    eo.add_configure(vecs=X, nlist=20)
    assert np.array_equal(eo.codes, eo.fine_quantizer.encode(X))
    ids, dists = eo.query(X[3], topk=5, method="linear")
    assert len(ids) == 5 and np.all(np.diff(dists) >= 0)
    # merge (tests/test_rii.py:252-289)
    e1 = rii.Rii(fine_quantizer=e.fine_quantizer)
    e1.add_configure(vecs=X[:400], nlist=5)
    e2 = rii.Rii(fine_quantizer=e.fine_quantizer)
    e2.add_configure(vecs=X[400:], nlist=5)
    e1.merge(e2)
    assert e1.N == N and np.array_equal(e1.codes, e.fine_quantizer.encode(X))
    assert sorted(sum(e1.posting_lists, [])) == list(range(N))


# ---------------------------------------------------------------- BASELINE sizes: properties ----------
def test_full_size_properties_c2_c3():
    """N = 1M, M = 32 (BASELINE C2/C3).  The oracle still finishes a handful of queries in seconds, so a few
    are checked bit-exactly; the rest through size-independent properties (ivf(L=N) == linear,
    subset(all ids) == no subset, subset results inside S, sortedness, idempotence)."""
    D, M, Ks, N, nlist = 128, 32, 256, 1000000, 1000
    cw, codes, Q = synth(D, M, Ks, N, 16, seed=2024)
    e = engine(cw, codes)
    # C2 index: coarse centers from a GPU reconfigure with iter=1; lists == oracle assignment
    e.reconfigure(nlist, 1)
    centers = e.coarse_centers_array()
    offsets, ids = e.posting_lists_csr()
    assert offsets[-1] == N and np.array_equal(np.sort(ids), np.arange(N))
    Dm = O.sym_matrices(cw)
    sl = slice(123456, 123456 + 20000)
    assert np.array_equal(O.assign(Dm, codes[sl], centers), _assign_of(offsets, ids, N)[sl])
    rng = np.random.default_rng(1)
    tids = np.sort(rng.choice(N, 100000, replace=False)).astype(np.int64)   # C3
    all_ids = np.arange(N, dtype=np.int64)
    for i, q in enumerate(Q):
        lin = e.query_linear(q, 10, EMPTY)
        assert lin == e.query_linear(q, 10, EMPTY)                           # idempotent
        assert all(lin[1][j] <= lin[1][j + 1] for j in range(9))
        assert e.query_linear(q, 10, all_ids) == lin
        assert e.query_ivf(q, 10, EMPTY, N) == lin                           # ivf over everything == linear
        sub = e.query_linear(q, 10, tids)
        assert set(sub[0]) <= set(tids.tolist())
        assert e.query_ivf(q, 10, tids, len(tids)) == sub
        ivf = e.query_ivf(q, 1, EMPTY, 32000)                                # C2 operating point
        assert len(ivf[0]) == 1
        if i < 3:  # bit-exact against the oracle at full size
            T = O.dtable(q, cw, 16)
            assert_same_result(lin[0], np.array(lin[1], np.float32), *O.query_linear(T, codes, 10), "C1/N=1M linear")
            assert_same_result(sub[0], np.array(sub[1], np.float32), *O.query_linear(T, codes, 10, tids), "C3")
            o = O.query_ivf(T, codes, centers, offsets, ids, 1, 32000)
            assert_same_result(ivf[0], np.array(ivf[1], np.float32), o[0], o[1], "C2")
            o = O.query_ivf(T, codes, centers, offsets, ids, 5, 32000, tids)
            g = e.query_ivf(q, 5, tids, 32000)
            assert_same_result(g[0], np.array(g[1], np.float32), o[0], o[1], "C2+C3")
    # batch entry == loop of single queries
    bi, bd, bc = e.query_batch(np.ascontiguousarray(Q), 1, L=32000, method="ivf")
    for i, q in enumerate(Q):
        one = e.query_ivf(q, 1, EMPTY, 32000)
        assert bi[i].tolist() == one[0] and bd[i].tolist() == one[1] and bc[i] == 1


def _assign_of(offsets, ids, N):
    a = np.empty(N, np.int32)
    for no in range(len(offsets) - 1):
        a[ids[offsets[no]:offsets[no + 1]]] = no
    return a


# ---------------------------------------------------------------- scan kernel v2 (skewed) --------------
# 4: k_scan_stream32 (two streams per lane, FFMA2, pre-skewed skew64 layout through private cp.async rings); 401 = always
# one CTA per SM (12 warps) instead of two CTAs of 6 warps for per-query IVF batches.  (The earlier skewed engines v2 / v3
# were removed in round 2.)
SKEW_KERNELS = [4, 401]


def set_kernel(e, sk):
    e.set_option("scan_kernel", 4 if sk > 100 else sk)
    e.set_option("stream_ctas", 1 if sk == 401 else 0)


@pytest.mark.parametrize("sk", SKEW_KERNELS)
@pytest.mark.parametrize("N", [1, 255, 2049, 20000, 70001, 300001])
def test_skew_kernel_matches_oracle(N, sk):
    """The bank-conflict-free schedules (k_scan_skew32, k_scan_dual32) must return exactly what the natural-layout
    kernel and the oracle return: every tail / tile-boundary / ring-wrap case."""
    D, M, Ks = 128, 32, 256
    cw, codes, Q = synth(D, M, Ks, N, 3, seed=N)
    e = engine(cw, codes)
    set_kernel(e, sk)
    for q in Q:
        T = O.dtable(q, cw, 16)
        for topk in (1, 10, 224):
            if topk > N:
                continue
            ids, d = e.query_linear(q, topk, EMPTY)
            assert_same_result(ids, np.array(d, np.float32), *O.query_linear(T, codes, topk), "skew N=%d k=%d" % (N, topk))
    if N >= 10:
        bi, bd, bc = e.query_batch(np.ascontiguousarray(Q), 10, method="linear")
        for b, q in enumerate(Q):
            exp = O.query_linear(O.dtable(q, cw, 16), codes, 10)
            assert_same_result(bi[b], bd[b], exp[0], exp[1], "skew batch")
    if N >= 10:  # forced streaming engine + unsorted target ids: refused (ascending ids run on a compact copy of the rows)
        with pytest.raises(Exception):
            e.query_linear(Q[0], 1, np.array([5, 2], dtype=np.int64))


@pytest.mark.parametrize("sk", SKEW_KERNELS)
def test_skew_kernel_small_ks_and_ties(sk):
    """Ks < 256 (table rows beyond Ks are never indexed by real codes) and massive exact ties."""
    cw, codes, Q = synth(128, 32, 16, 50000, 2, seed=3)
    codes[:, 8:] = 0  # only 16^8 distinct codes -> many exact distance ties; (dist, id) order decides
    e1, e2 = engine(cw, codes), engine(cw, codes)
    e1.set_option("scan_kernel", 1)
    set_kernel(e2, sk)
    for q in Q:
        T = O.dtable(q, cw, 16)
        for topk in (1, 50, 200):
            r1, r2 = e1.query_linear(q, topk, EMPTY), e2.query_linear(q, topk, EMPTY)
            assert r1 == r2
            assert_same_result(r2[0], np.array(r2[1], np.float32), *O.query_linear(T, codes, topk), "ties")


@pytest.mark.parametrize("N", [32767, 32768, 40000])
def test_linear_auto_dispatch_boundary(N):
    """The automatic choice between the natural-layout kernels (N < 32768) and the streaming engine (one launch: table, scan
    and the merge by the last CTA) must not change a bit of the result on either side of the boundary."""
    cw, codes, Q = synth(128, 32, 256, N, 3, seed=N)
    e, e1 = engine(cw, codes), engine(cw, codes)
    e1.set_option("scan_kernel", 1)
    for q in Q:
        T = O.dtable(q, cw, 16)
        for topk in (1, 10, 100):
            r, r1 = e.query_linear(q, topk, EMPTY), e1.query_linear(q, topk, EMPTY)
            assert r == r1
            assert_same_result(r[0], np.array(r[1], np.float32), *O.query_linear(T, codes, topk), "boundary N=%d k=%d" % (N, topk))
    bi, bd, bc = e.query_batch(Q, 7, method="linear")
    for b, q in enumerate(Q):
        assert_same_result(bi[b], bd[b], *O.query_linear(O.dtable(q, cw, 16), codes, 7), "boundary batch")


def test_skew_kernel_large_n_vs_v1():
    N = 5000000
    cw, codes, Q = synth(128, 32, 256, N, 3, seed=77)
    e = engine(cw, codes)
    for q in Q:
        e.set_option("scan_kernel", 1)
        r1 = e.query_linear(q, 100, EMPTY)
        e.set_option("scan_kernel", 0)  # auto -> v4 at this size
        r0 = e.query_linear(q, 100, EMPTY)
        e.set_option("scan_kernel", 4)
        r2 = e.query_linear(q, 100, EMPTY)
        assert r1 == r2 == r0
    T = O.dtable(Q[0], cw, 16)
    assert_same_result(r2[0], np.array(r2[1], np.float32), *O.query_linear(O.dtable(Q[2], cw, 16), codes, 100), "5M")


@pytest.mark.parametrize("sk", SKEW_KERNELS)
def test_ivf_skew_kernel_vs_v1_and_oracle(sk):
    """The skewed posting-list scans (auto-selected for M = 32, no target_ids, topk <= 224) against the natural-layout
    kernel and the oracle: single-candidate plans, mid-list cuts, several CTAs per query, batches."""
    D, M, Ks, N, nlist = 128, 32, 256, 60000, 50
    cw, codes, Q = synth(D, M, Ks, N, 5, seed=31)
    e = engine(cw, codes)
    e.reconfigure(nlist, 2)
    centers = e.coarse_centers_array()
    offsets, ids = e.posting_lists_csr()
    cases = [(1, 1), (1, 7), (3, 255), (1, 256), (10, 257), (96, 96), (5, 1200), (1, 2048), (20, 2049), (50, 33333), (96, N)]
    for q in Q:
        T = O.dtable(q, cw, 16)
        for topk, L in cases:
            exp = O.query_ivf(T, codes, centers, offsets, ids, topk, L)
            set_kernel(e, sk)
            r2 = e.query_ivf(q, topk, EMPTY, L)
            e.set_option("scan_kernel", 1)
            r1 = e.query_ivf(q, topk, EMPTY, L)
            assert r1 == r2, (topk, L)
            assert_same_result(r2[0], np.array(r2[1], np.float32), exp[0], exp[1], "ivf skew k=%d L=%d" % (topk, L))
    set_kernel(e, sk)
    Qb = np.ascontiguousarray(np.tile(Q, (60, 1)))  # 300 queries -> one CTA per query
    bi, bd, bc = e.query_batch(Qb, 4, L=3000, method="ivf")
    for b in range(0, 300, 37):
        exp = O.query_ivf(O.dtable(Qb[b], cw, 16), codes, centers, offsets, ids, 4, 3000)
        assert_same_result(bi[b], bd[b], exp[0], exp[1], "ivf skew batch %d" % b)


def test_gpu_pq_encoder():
    """rii_encode: nearest codeword per sub-space, fp32 sequential sum, first minimum wins --
    exact against a numpy float32 restatement, and equal to the host codec (scipy vq) except on near-ties."""
    import rii_b200 as rii
    rng = np.random.default_rng(5)
    for D, M, Ks in [(128, 32, 256), (40, 4, 20), (96, 32, 256), (80, 4, 16)]:
        Ds = D // M
        X = rng.random((3000, D), dtype=np.float32)
        codec = rii.PQ(M=M, Ks=Ks, verbose=False).fit(X[:1500], iter=5)
        e = engine(codec.codewords)
        got = e.encode(X)
        exp = np.empty_like(got)
        for m in range(M):
            sub = X[:, m * Ds:(m + 1) * Ds]
            d = np.zeros((X.shape[0], Ks), np.float32)
            for i in range(Ds):
                t = (sub[:, i:i + 1] - codec.codewords[m][None, :, i]).astype(np.float32)
                d = (d + (t * t).astype(np.float32)).astype(np.float32)
            exp[:, m] = d.argmin(1)
        assert np.array_equal(got, exp)
        assert (got == codec.encode(X)).mean() > 0.999
    e2 = rii.Rii(fine_quantizer=codec)
    e2.add(X, gpu_encode=True)
    assert np.array_equal(e2.codes, got)


@pytest.mark.parametrize("sk", SKEW_KERNELS)
def test_fused_coarse_plan_scan_equals_unfused(sk):
    """One-CTA-per-query batches run coarse ranking + plan + posting-list scan in ONE kernel (two passes of the
    skewed engine); results must equal the unfused pipeline (k_coarse_rank + scan) and the oracle,
    including plans that end mid-list, at the w-th list, and the empty result."""
    D, M, Ks, N, nlist = 128, 32, 256, 80000, 200
    cw, codes, Q = synth(D, M, Ks, N, 160, seed=77)
    e = engine(cw, codes)
    set_kernel(e, sk)
    e.reconfigure(nlist, 1)
    centers = e.coarse_centers_array()
    offsets, ids = e.posting_lists_csr()
    Qb = np.ascontiguousarray(Q)
    for topk, L in [(1, 400), (7, 3000), (96, 97), (3, N), (20, 12345), (1, 1)]:
        e.set_option("fuse_coarse", 1)
        f = e.query_batch(Qb, topk, L=L, method="ivf")
        e.set_option("fuse_coarse", 0)
        u = e.query_batch(Qb, topk, L=L, method="ivf")
        assert np.array_equal(f[2], u[2]) and np.array_equal(f[0], u[0]) and np.array_equal(bits(f[1]), bits(u[1])), (topk, L)
        for b in range(0, 160, 23):
            exp = O.query_ivf(O.dtable(Qb[b], cw, 16), codes, centers, offsets, ids, topk, L)
            n = int(f[2][b])
            assert_same_result(f[0][b][:n], f[1][b][:n], exp[0], exp[1], "fused k=%d L=%d b=%d" % (topk, L, b))
    e.set_option("fuse_coarse", 1)


@pytest.mark.parametrize("sk", [4])
def test_dual_kernel_huge_tables_take_the_plain_path(sk):
    """k_scan_dual32 / k_scan_stream32 accumulate with acc * {0,1} + v, which needs finite partial sums; a distance table with huge /
    inf entries (queries ~1e19 away from the codewords) must be detected in-kernel and scanned by the plain
    per-candidate loop -- same results as the natural-layout kernel and the oracle, bit for bit."""
    D, M, Ks, N, nlist = 128, 32, 256, 40000, 40
    cw, codes, Q = synth(D, M, Ks, N, 4, seed=9)
    Q = Q.copy()
    Q[0, 5] = 4e18      # entries ~1.6e37 > 1e37: plain path, finite sums
    Q[1, 17] = 3e19     # (3e19)^2 overflows: inf entries in one sub-space -> every distance inf, ids decide
    Q[2, :] = 2.5e18    # every sub-space huge: sums reach 1e38, some overflow to inf
    e = engine(cw, codes)
    e.reconfigure(nlist, 1)
    centers = e.coarse_centers_array()
    offsets, ids = e.posting_lists_csr()
    for q in Q:
        T = O.dtable(q, cw, 16)
        for topk in (1, 20):
            e.set_option("scan_kernel", sk)
            r3 = e.query_linear(q, topk, EMPTY)
            e.set_option("scan_kernel", 1)
            r1 = e.query_linear(q, topk, EMPTY)
            assert r1[0] == r3[0] and bits(np.array(r1[1], np.float32)).tolist() == bits(np.array(r3[1], np.float32)).tolist()
            assert_same_result(r3[0], np.array(r3[1], np.float32), *O.query_linear(T, codes, topk), "huge linear")
            exp = O.query_ivf(T, codes, centers, offsets, ids, topk, 3000)
            e.set_option("scan_kernel", sk)
            g = e.query_ivf(q, topk, EMPTY, 3000)
            assert_same_result(g[0], np.array(g[1], np.float32), exp[0], exp[1], "huge ivf")
    e.set_option("scan_kernel", sk)
    bi, bd, bc = e.query_batch(np.ascontiguousarray(np.tile(Q, (40, 1))), 3, L=2000, method="ivf")  # fused, one CTA per query
    for b in range(4):
        exp = O.query_ivf(O.dtable(Q[b], cw, 16), codes, centers, offsets, ids, 3, 2000)
        assert_same_result(bi[b], bd[b], exp[0], exp[1], "huge fused %d" % b)


def test_stream_kernel_tracks_index_updates():
    """The skew64 copies are derived state: add() after a query, add with posting-list update, reconfigure and clear
    must all be reflected by the next query (rebuilt lazily)."""
    D, M, Ks = 128, 32, 256
    cw, codes, Q = synth(D, M, Ks, 9000, 3, seed=12)
    e = engine(cw, codes[:3000])
    e.set_option("scan_kernel", 4)
    q = Q[0]
    T = O.dtable(q, cw, 16)
    r = e.query_linear(q, 5, EMPTY)
    assert_same_result(r[0], np.array(r[1], np.float32), *O.query_linear(T, codes[:3000], 5), "before add")
    e.add_codes(codes[3000:5001], False)
    r = e.query_linear(q, 5, EMPTY)
    assert_same_result(r[0], np.array(r[1], np.float32), *O.query_linear(T, codes[:5001], 5), "after add")
    e.reconfigure(30, 2)
    e.add_codes(codes[5001:], True)
    centers = e.coarse_centers_array()
    offsets, ids = e.posting_lists_csr()
    for topk, L in [(1, 100), (5, 999), (9, 9000)]:
        exp = O.query_ivf(T, codes, centers, offsets, ids, topk, L)
        g = e.query_ivf(q, topk, EMPTY, L)
        assert_same_result(g[0], np.array(g[1], np.float32), exp[0], exp[1], "ivf after add+update k=%d L=%d" % (topk, L))
    e.reconfigure(7, 1)
    centers = e.coarse_centers_array()
    offsets, ids = e.posting_lists_csr()
    bi, bd, bc = e.query_batch(np.ascontiguousarray(np.tile(Q, (50, 1))), 2, L=500, method="ivf")  # fused
    for b in range(3):
        exp = O.query_ivf(O.dtable(Q[b], cw, 16), codes, centers, offsets, ids, 2, 500)
        assert_same_result(bi[b], bd[b], exp[0], exp[1], "after second reconfigure %d" % b)


@pytest.mark.parametrize("sk", [4, 401])
def test_stream_fused_large_nlist(sk):
    """nlist > 1024: the fused v4 kernel ranks the coarse centers in the warps' top-k lists (pass 0 is an ordinary
    scan with k = w) instead of the histogram select; same results as the unfused v1 pipeline and the oracle."""
    D, M, Ks, N, nlist = 128, 32, 256, 150000, 1500
    cw, codes, Q = synth(D, M, Ks, N, 150, seed=5)
    e = engine(cw, codes)
    e.reconfigure(nlist, 1)
    centers = e.coarse_centers_array()
    offsets, ids = e.posting_lists_csr()
    Qb = np.ascontiguousarray(Q)
    for topk, L in [(1, 300), (5, 4000), (40, 100), (3, 20000)]:   # w = 6, 43, 4, 203 (<= 224: the lists hold them)
        set_kernel(e, sk)
        f = e.query_batch(Qb, topk, L=L, method="ivf")
        e.set_option("scan_kernel", 1)
        u = e.query_batch(Qb, topk, L=L, method="ivf")
        assert np.array_equal(f[2], u[2]) and np.array_equal(f[0], u[0]) and np.array_equal(bits(f[1]), bits(u[1])), (topk, L)
        for b in range(0, 150, 31):
            exp = O.query_ivf(O.dtable(Qb[b], cw, 16), codes, centers, offsets, ids, topk, L)
            n = int(f[2][b])
            assert_same_result(f[0][b][:n], f[1][b][:n], exp[0], exp[1], "large nlist k=%d L=%d b=%d" % (topk, L, b))


@pytest.mark.parametrize("N", [1, 63, 64, 65, 1000, 40001, 300001])
def test_stream_kernel_m64_linear(N):
    """M = 64 on the v4 engine (two 64 KB tables, a row = two 32-byte half-row blocks): every tail / group-boundary
    case against the natural-layout kernel and the oracle."""
    D, M, Ks = 128, 64, 256
    cw, codes, Q = synth(D, M, Ks, N, 3, seed=100 + N)
    e = engine(cw, codes)
    for q in Q:
        T = O.dtable(q, cw, 16)
        for topk in (1, 10, 224):
            if topk > N:
                continue
            e.set_option("scan_kernel", 4)
            r4 = e.query_linear(q, topk, EMPTY)
            e.set_option("scan_kernel", 1)
            r1 = e.query_linear(q, topk, EMPTY)
            assert r1 == r4, (N, topk)
            assert_same_result(r4[0], np.array(r4[1], np.float32), *O.query_linear(T, codes, topk), "m64 N=%d k=%d" % (N, topk))
    if N >= 10:
        e.set_option("scan_kernel", 4)
        bi, bd, bc = e.query_batch(np.ascontiguousarray(Q), 10, method="linear")
        for b, q in enumerate(Q):
            exp = O.query_linear(O.dtable(q, cw, 16), codes, 10)
            assert_same_result(bi[b], bd[b], exp[0], exp[1], "m64 batch")


def test_stream_kernel_m64_ivf():
    """M = 64 posting-list scans on the v4 engine: unfused plans (few queries, several CTAs per query), fused
    one-CTA-per-query batches with nlist <= 1024 (histogram select) and > 1024 (warp lists), Ds = 2 and Ds = 3."""
    for D, nlist, N in [(128, 60, 70000), (192, 1100, 120000)]:
        M, Ks = 64, 256
        cw, codes, Q = synth(D, M, Ks, N, 150, seed=D)
        e = engine(cw, codes)
        e.reconfigure(nlist, 1)
        centers = e.coarse_centers_array()
        offsets, ids = e.posting_lists_csr()
        for q in Q[:3]:
            T = O.dtable(q, cw, 16)
            for topk, L in [(1, 1), (3, 255), (10, 2049), (50, 33333), (96, N)]:
                exp = O.query_ivf(T, codes, centers, offsets, ids, topk, L)
                e.set_option("scan_kernel", 4)
                try:
                    r4 = e.query_ivf(q, topk, EMPTY, L)
                except Exception:  # w > 224 lists: not a v4 shape when forced
                    e.set_option("scan_kernel", 0)
                    r4 = e.query_ivf(q, topk, EMPTY, L)
                assert_same_result(r4[0], np.array(r4[1], np.float32), exp[0], exp[1], "m64 ivf D=%d k=%d L=%d" % (D, topk, L))
        Qb = np.ascontiguousarray(Q)
        for topk, L in [(1, 400), (7, 3000), (20, 12345)]:
            e.set_option("scan_kernel", 4)
            f = e.query_batch(Qb, topk, L=L, method="ivf")
            e.set_option("scan_kernel", 1)
            u = e.query_batch(Qb, topk, L=L, method="ivf")
            assert np.array_equal(f[2], u[2]) and np.array_equal(f[0], u[0]) and np.array_equal(bits(f[1]), bits(u[1])), (D, topk, L)
            for b in range(0, 150, 37):
                exp = O.query_ivf(O.dtable(Qb[b], cw, 16), codes, centers, offsets, ids, topk, L)
                n = int(f[2][b])
                assert_same_result(f[0][b][:n], f[1][b][:n], exp[0], exp[1], "m64 fused D=%d k=%d L=%d b=%d" % (D, topk, L, b))


def test_two_phase_subset_api_single_shard():
    """rii_ivf_subset_counts_dev / rii_ivf_subset_scan_dev (the sharded IVF + target_ids path) with one shard must equal
    the ordinary query_ivf(target_ids): uniform subset, subset concentrated in far lists (flagged -> full ranking)."""
    import torch
    from rii_b200 import sharded
    D, M, Ks, N, nlist = 128, 32, 256, 60000, 60
    cw, codes, Q = synth(D, M, Ks, N, 6, seed=77)
    e = engine(cw, codes)
    e.reconfigure(nlist, 1)
    centers = e.coarse_centers_array()
    offsets, ids = e.posting_lists_csr()
    eng = sharded.CudaShardEngine(e)
    dev = torch.device("cuda", 0)
    Qd = torch.from_numpy(Q).to(dev)
    rng = np.random.default_rng(3)
    far = np.argsort(O.adist_all(O.dtable(Q[0], cw, 16), centers))[-4:]
    far_ids = np.sort(np.concatenate([ids[offsets[no]:offsets[no + 1]] for no in far])).astype(np.int64)
    for tids, topk, L in [(np.sort(rng.choice(N, 6000, replace=False)).astype(np.int64), 10, 1000), (far_ids, 3, 30)]:
        gi, gd, gc = sharded.sharded_query_subset(eng, Qd, topk, L, torch.from_numpy(tids).to(dev), None, 1, 0)
        torch.cuda.synchronize()
        gi, gd, gc = gi.cpu().numpy(), gd.cpu().numpy(), gc.cpu().numpy()
        for b, q in enumerate(Q):
            exp = O.query_ivf(O.dtable(q, cw, 16), codes, centers, offsets, ids, topk, L, tids)
            n = int(gc[b])
            assert_same_result(gi[b, :n], gd[b, :n], exp[0], exp[1], "two-phase subset k=%d L=%d b=%d" % (topk, L, b))
            one = e.query_ivf(q, topk, tids, L)
            assert one[0] == gi[b, :n].tolist()


@pytest.mark.parametrize("M", [32, 64])
def test_stream_kernel_incremental_adds(M):
    """A stream of add() calls interleaved with linear queries: the skew64 copy is extended from the first incomplete
    group of 64 rows on (not rebuilt), and every query must see exactly the codes added so far."""
    D, Ks = 128, 256
    cw, codes, Q = synth(D, M, Ks, 5000, 2, seed=M)
    e = engine(cw)
    e.set_option("scan_kernel", 4)
    n = 0
    for step in (1, 62, 1, 1, 64, 127, 500, 3, 1000, 2241, 1000):
        e.add_codes(codes[n:n + step], False)
        n += step
        for q in Q:
            k = min(5, n)
            r = e.query_linear(q, k, EMPTY)
            assert_same_result(r[0], np.array(r[1], np.float32), *O.query_linear(O.dtable(q, cw, 16), codes[:n], k), "after %d rows" % n)
    assert n == 5000


@pytest.mark.parametrize("G", [2, 3])
def test_id_range_shards_on_one_gpu(G):
    """SURVEY 8e on ONE device: G id-range shards as G handles on cuda:0 (rii_set_shard, rii_fit_coarse,
    rii_set_coarse_centers, rii_set_global_lengths; concat of the per-shard outputs + rii_merge_shards_dev) must equal the
    unsharded oracle bit for bit: linear, IVF (the global cut planned from glob_len / pre_len) and IVF + target_ids."""
    import torch
    from rii_b200 import sharded
    D, M, Ks, N, nlist = 128, 32, 256, 150000, 90
    cw, codes, Q = synth(D, M, Ks, N, 6, seed=4321)
    grp = sharded.LocalShardGroup([engine(cw) for _ in range(G)])
    centers = grp.build(codes, nlist, 2)
    oc, oa = O.reconfigure(cw, codes, nlist, 2)
    assert np.array_equal(centers, oc), "sharded coarse centers differ from the oracle"
    offsets, ids = O.assign_to_lists(oa, nlist)
    # the shards' posting lists are the id-range slices of the oracle's lists
    b = sharded.shard_bounds(N, G)
    for g, eng in enumerate(grp.engines):
        so, si = eng.e.posting_lists_csr()
        for no in range(0, nlist, 7):
            full = ids[offsets[no]:offsets[no + 1]]
            assert np.array_equal(si[so[no]:so[no + 1]] + b[g], full[(full >= b[g]) & (full < b[g + 1])])
    dev = torch.device("cuda", 0)
    Qd = torch.from_numpy(Q).to(dev)
    for method, topk, L in [("linear", 1, 0), ("linear", 50, 0), ("ivf", 1, 2000), ("ivf", 10, 6400), ("ivf", 100, 150),
                            ("ivf", 5, N), ("ivf", 3, 1666)]:
        runs = [("", grp.query(Qd, topk, L, method))]
        if method == "ivf":  # the coarse phase split over the shards' handles (rii_coarse_rank_dev / rii_query_ranked_dev)
            runs.append(("split ", grp.query_split(Qd, topk, L)))
        for tag, (gi, gd, gc) in runs:
            torch.cuda.synchronize()
            gi, gd, gc = gi.cpu().numpy(), gd.cpu().numpy(), gc.cpu().numpy()
            for bq, q in enumerate(Q):
                T = O.dtable(q, cw, 16)
                exp = O.query_linear(T, codes, topk) if method == "linear" else O.query_ivf(T, codes, oc, offsets, ids, topk, L)
                n = int(gc[bq])
                assert_same_result(gi[bq, :n], gd[bq, :n], exp[0], exp[1], "shards G=%d %s%s k=%d L=%d" % (G, tag, method, topk, L))
    rng = np.random.default_rng(5)
    far = np.argsort(O.adist_all(O.dtable(Q[0], cw, 16), oc))[-5:]
    far_ids = np.sort(np.concatenate([ids[offsets[no]:offsets[no + 1]] for no in far])).astype(np.int64)
    for tids, topk, L in [(np.sort(rng.choice(N, 20000, replace=False)).astype(np.int64), 10, 3200),
                          (np.sort(rng.choice(N, 3000, replace=False)).astype(np.int64), 1, 3000), (far_ids, 3, 40)]:
        gi, gd, gc = grp.query_subset(Qd, topk, L, torch.from_numpy(tids).to(dev))
        torch.cuda.synchronize()
        gi, gd, gc = gi.cpu().numpy(), gd.cpu().numpy(), gc.cpu().numpy()
        for bq, q in enumerate(Q):
            exp = O.query_ivf(O.dtable(q, cw, 16), codes, oc, offsets, ids, topk, L, tids)
            n = int(gc[bq])
            assert_same_result(gi[bq, :n], gd[bq, :n], exp[0], exp[1], "shards G=%d subset k=%d L=%d" % (G, topk, L))
    # mutating a shard invalidates its global state until the lengths are exchanged again
    grp.engines[0].e.add_codes(codes[:10], False)
    with pytest.raises(Exception):
        grp.engines[0].query_local(Qd, 1, 1000, "ivf")


@pytest.mark.parametrize("D,M,Ks", [(40, 20, 256), (96, 24, 64), (96, 48, 256), (60, 12, 256)])
def test_padded_rows_on_the_streaming_engine(D, M, Ks):
    """12 <= M <= 64 that are not 32 / 64 run on the streaming engine with zero-padded rows and zero table columns
    (x + 0.0f == x: still the reference's sequential sum): linear, IVF, and the assignment engine."""
    N, nlist = 30000, 40
    cw, codes, Q = synth(D, M, Ks, N, 4, seed=M * 7)
    e = engine(cw, codes)
    e.reconfigure(nlist, 2)
    oc, oa = O.reconfigure(cw, codes, nlist, 2)
    assert np.array_equal(e.coarse_centers_array(), oc)
    offsets, ids = O.assign_to_lists(oa, nlist)
    eo, ei = e.posting_lists_csr()
    assert np.array_equal(eo, offsets) and np.array_equal(ei, ids)
    e.set_option("scan_kernel", 4)
    for q in Q:
        T = O.dtable(q, cw, 16)
        for topk in (1, 17):
            r = e.query_linear(q, topk, EMPTY)
            assert_same_result(r[0], np.array(r[1], np.float32), *O.query_linear(T, codes, topk), "padded linear M=%d" % M)
            for L in (topk, 900, 5000):
                r = e.query_ivf(q, topk, EMPTY, L)
                exp = O.query_ivf(T, codes, oc, offsets, ids, topk, L)
                assert_same_result(r[0], np.array(r[1], np.float32), exp[0], exp[1], "padded ivf M=%d L=%d" % (M, L))


@pytest.mark.parametrize("M,N,K", [(32, 1, 1), (32, 63, 5), (32, 64, 64), (32, 65, 300), (32, 5000, 1000), (32, 100000, 37),
                                   (64, 777, 129), (64, 40000, 50), (20, 3000, 77), (48, 3000, 77)])
def test_assign_stream_engine_vs_natural_and_oracle(M, N, K):
    """K6 on the streaming engine (k_assign_stream + k_assign_reduce) == natural-layout k_assign == oracle: assignments
    and distances bit for bit, every row / center tail, first minimum wins (duplicated centers)."""
    D, Ks = 4 * M, 256
    cw, codes, _ = synth(D, M, Ks, N, 1, seed=N + K)
    rng = np.random.default_rng(N * 31 + K)
    centers = rng.integers(0, Ks, (K, M), dtype=np.uint8)
    if K > 4:
        centers[K // 2] = centers[1]  # exact tie between two centers: the lower index must win
        centers[K - 1] = centers[0]
    e = engine(cw)
    res = {}
    for ak in (1, 0, 3):
        e.set_option("assign_kernel", ak)
        res[ak] = e.assign(codes, centers, return_dist=True)
    for ak in (0, 3):
        assert np.array_equal(res[ak][0], res[1][0]), "assign_kernel=%d" % ak
        assert np.array_equal(bits(res[ak][1]), bits(res[1][1])), "assign_kernel=%d" % ak
    n_chk = min(N, 2000)
    Dm = O.sym_matrices(cw)
    oa, od = O.assign(Dm, codes[:n_chk], centers, return_dist=True)
    assert np.array_equal(res[0][0][:n_chk], oa) and np.array_equal(bits(res[0][1][:n_chk]), bits(od))


@pytest.mark.parametrize("M", [32, 64, 20])
def test_subset_search_on_the_streaming_engine(M):
    """target_ids on the streaming engine (src/rii.h:218-228,294): linear over a compact skew64 copy of
    the target rows, IVF over the sub-index of the members -- against the oracle and the natural-layout kernels; sparse
    and dense subsets, repeated ids, many ranked lists (w > 224), subsets concentrated in far lists (walk beyond w), and
    the empty result."""
    D, Ks, N, nlist = 4 * M, 256, 60000, 300
    cw, codes, Q = synth(D, M, Ks, N, 5, seed=500 + M)
    e = engine(cw, codes)
    e.reconfigure(nlist, 1)
    centers = e.coarse_centers_array()
    offsets, ids = e.posting_lists_csr()
    rng = np.random.default_rng(11)
    far = np.argsort(O.adist_all(O.dtable(Q[0], cw, 16), centers))[-6:]
    far_ids = np.sort(np.concatenate([ids[offsets[no]:offsets[no + 1]] for no in far])).astype(np.int64)
    dup = np.sort(np.concatenate([rng.choice(N, 3000, replace=False), np.arange(100, 140)])).astype(np.int64)
    dup = np.sort(np.concatenate([dup, dup[:500]]))
    cases = [(np.sort(rng.choice(N, 6000, replace=False)).astype(np.int64), 10, 1000),      # w = 53
             (np.sort(rng.choice(N, 6000, replace=False)).astype(np.int64), 3, 5000),       # w = 253 > 224 ranked lists
             (np.sort(rng.choice(N, 30000, replace=False)).astype(np.int64), 1, 30000),     # L = S: every member
             (np.sort(rng.choice(N, 200, replace=False)).astype(np.int64), 5, 150),
             (dup, 7, 800), (far_ids, 3, 30), (far_ids[:5], 5, 5), (np.arange(N, dtype=np.int64), 10, 2000)]
    for tids, topk, L in cases:
        for q in Q[:3]:
            T = O.dtable(q, cw, 16)
            exp = O.query_ivf(T, codes, centers, offsets, ids, topk, L, tids)
            for sk in (4, 1):
                e.set_option("scan_kernel", sk)
                try:
                    g = e.query_ivf(q, topk, tids, L)
                except Exception:
                    if sk == 4 and int(round(L * nlist / len(tids))) + 3 >= nlist:
                        continue  # the forced engine cannot re-rank every list when nlist is large; auto falls back
                    raise
                assert_same_result(g[0], np.array(g[1], np.float32), exp[0], exp[1], "subset ivf M=%d sk=%d S=%d k=%d L=%d" % (M, sk, len(tids), topk, L))
            expl = O.query_linear(T, codes, topk, tids)
            for sk in (4, 1):
                e.set_option("scan_kernel", sk)
                g = e.query_linear(q, topk, tids)
                if len(np.unique(tids)) == len(tids):
                    assert_same_result(g[0], np.array(g[1], np.float32), expl[0], expl[1], "subset linear M=%d sk=%d S=%d k=%d" % (M, sk, len(tids), topk))
                else:  # repeated ids give repeated results (order among equal keys is free): compare as multisets of keys
                    assert sorted(zip(np.array(g[1], np.float32).tolist(), g[0])) == sorted(zip(expl[1].tolist(), expl[0].tolist()))
    # unsorted ids (linear only; src/rii.h:218-228 takes them in the given order): natural-layout kernel, same result set
    e.set_option("scan_kernel", 0)
    t = rng.permutation(N)[:5000].astype(np.int64)
    g = e.query_linear(Q[0], 20, t)
    exp = O.query_linear(O.dtable(Q[0], cw, 16), codes, 20, t)
    assert_same_result(g[0], np.array(g[1], np.float32), exp[0], exp[1], "unsorted subset")
    # batch entry: one sub-index serves the whole batch
    tids = cases[0][0]
    Qb = np.ascontiguousarray(np.tile(Q, (40, 1)))
    bi, bd, bc = e.query_batch(Qb, 4, target_ids=tids, L=1000, method="ivf")
    li, ld, lc = e.query_batch(Qb, 4, target_ids=tids, method="linear")
    for b in range(0, 200, 23):
        T = O.dtable(Qb[b], cw, 16)
        exp = O.query_ivf(T, codes, centers, offsets, ids, 4, 1000, tids)
        assert_same_result(bi[b][:bc[b]], bd[b][:bc[b]], exp[0], exp[1], "subset batch ivf")
        exp = O.query_linear(T, codes, 4, tids)
        assert_same_result(li[b][:lc[b]], ld[b][:lc[b]], exp[0], exp[1], "subset batch linear")


@pytest.mark.parametrize("M,Ks", [(32, 256), (8, 64), (72, 256)])
def test_any_topk_and_any_number_of_ranked_lists(M, Ks):
    """The reference allows 1 <= topk <= N (rii/rii.py:280-281: topk=None means N) and re-ranks ALL lists when the first
    w hold too few candidates; shapes whose keys do not fit shared memory take the global-memory path (keys in HBM + radix
    sort).  ADVICE r1: large topk, and nlist >= 8192."""
    D, N = 4 * M, 40000
    cw, codes, Q = synth(D, M, Ks, N, 2, seed=900 + M)
    e = engine(cw, codes)
    for q in Q:
        T = O.dtable(q, cw, 16)
        for topk in (300, 20000, N):
            g = e.query_linear(q, topk, EMPTY)
            assert_same_result(g[0], np.array(g[1], np.float32), *O.query_linear(T, codes, topk), "large topk linear M=%d k=%d" % (M, topk))
        tids = np.arange(0, N, 3, dtype=np.int64)
        g = e.query_linear(q, len(tids), tids)
        assert_same_result(g[0], np.array(g[1], np.float32), *O.query_linear(T, codes, len(tids), tids), "large topk subset")
    nlist = 9000 if M == 32 else 500
    e.reconfigure(nlist, 1)
    centers = e.coarse_centers_array()
    offsets, ids = e.posting_lists_csr()
    for q in Q:
        T = O.dtable(q, cw, 16)
        for topk, L in [(300, 4000), (5000, 5000), (N, N), (2, 10)]:
            g = e.query_ivf(q, topk, EMPTY, L)
            exp = O.query_ivf(T, codes, centers, offsets, ids, topk, L)
            assert_same_result(g[0], np.array(g[1], np.float32), exp[0], exp[1], "large topk ivf M=%d k=%d L=%d" % (M, topk, L))
        # a subset concentrated in the farthest lists: fewer than topk members in the first w lists -> every list is ranked
        far = np.argsort(O.adist_all(T, centers))[-8:]
        tids = np.sort(np.concatenate([ids[offsets[no]:offsets[no + 1]] for no in far])).astype(np.int64)
        if len(tids) >= 3:
            g = e.query_ivf(q, 3, tids, min(len(tids), 12))
            exp = O.query_ivf(T, codes, centers, offsets, ids, 3, min(len(tids), 12), tids)
            assert_same_result(g[0], np.array(g[1], np.float32), exp[0], exp[1], "full re-ranking M=%d nlist=%d" % (M, nlist))


def test_reconfigure_with_large_nlist_sets_a_threshold():
    """Rii.reconfigure() with nlist >= 8192 must finish with a usable threshold function."""
    from rii_b200 import Rii, pq
    rng = np.random.default_rng(1)
    D, M, N = 64, 16, 70000
    X = rng.random((N, D), dtype=np.float32)
    codec = pq.PQ(M=M, Ks=256, verbose=False).fit(X[:3000], iter=3, seed=123)
    e = Rii(codec)
    e.add_configure(X, nlist=8200, iter=1)
    assert e.threshold is not None and e.nlist == 8200
    q = X[17]
    ids, d = e.query(q, topk=3)
    ids2, d2 = e.query(q, topk=3, method="linear")
    assert ids2[0] == ids[0] or d[0] >= d2[0]


def test_state_exchange_formats(tmp_path):
    """(1) the flat memory-mappable format round-trips an index bit for bit; (2) the reference's own
    5-tuple pickle state (src/main.cpp:35-54) exported from the GPU index loads in the unmodified reference (when its build
    travelled with the repo) and answers like us."""
    import os
    import pickle
    import subprocess
    import sys
    from rii_b200 import main
    from oracle import ref as R
    cw, codes, Q = synth(64, 16, 256, 5000, 3, seed=8)
    e = engine(cw, codes)
    e.reconfigure(30, 2)
    d = str(tmp_path / "flat")
    e.save_flat(d)
    e2 = main.RiiCpp.load_flat(d)
    assert e2.N == e.N and e2.nlist == e.nlist
    assert np.array_equal(e2.codes_array(), codes) and np.array_equal(e2.coarse_centers_array(), e.coarse_centers_array())
    o1, i1 = e.posting_lists_csr()
    o2, i2 = e2.posting_lists_csr()
    assert np.array_equal(o1, o2) and np.array_equal(i1, i2)
    tids = np.arange(0, 5000, 7, dtype=np.int64)
    for q in Q:
        assert e.query_ivf(q, 5, EMPTY, 700) == e2.query_ivf(q, 5, EMPTY, 700)
        assert e.query_ivf(q, 5, tids, 300) == e2.query_ivf(q, 5, tids, 300)  # (the row -> list map is rebuilt on load)
    # our own pickle (arrays) still round-trips
    e3 = pickle.loads(pickle.dumps(e))
    assert e3.query_linear(Q[0], 4, EMPTY) == e.query_linear(Q[0], 4, EMPTY)
    st = e.to_reference_state()
    assert len(st) == 5 and isinstance(st[3], list) and len(st[3]) == 5000 * 16 and isinstance(st[4][0], list)
    so_dir = R.variant_dir("strict")
    if so_dir is None:
        return
    f = str(tmp_path / "ref.pkl")
    open(f, "wb").write(e.dumps_reference())
    np.save(str(tmp_path / "q.npy"), Q[0])
    code = ("import sys, pickle, numpy as np; sys.path.insert(0, %r); import main; e = pickle.load(open(%r, 'rb'));"
            "q = np.load(%r); print(e.query_ivf(q, 3, np.array([], np.int64), 700)[0])" % (so_dir, f, str(tmp_path / "q.npy")))
    out = subprocess.check_output([sys.executable, "-c", code], stderr=subprocess.STDOUT).decode()
    assert str(e.query_ivf(Q[0], 3, EMPTY, 700)[0]) in out, out


@pytest.mark.parametrize("M,nlist", [(32, 1500), (64, 1100)])
def test_split_coarse_phase_with_many_lists(M, nlist):
    """nlist > 1024: the coarse pass ranks the centers in the warps' top-k lists.  Coarse-only launch
    (rii_coarse_rank_dev) + scan with the given ranking (rii_query_ranked_dev) over G = 2 id-range shards == oracle."""
    import torch
    from rii_b200 import sharded
    D, Ks, N = 4 * M, 256, 120000
    cw, codes, Q = synth(D, M, Ks, N, 8, seed=77 + M)
    grp = sharded.LocalShardGroup([engine(cw) for _ in range(2)])
    centers = grp.build(codes, nlist, 1)
    oc, oa = O.reconfigure(cw, codes, nlist, 1)
    assert np.array_equal(centers, oc)
    offsets, ids = O.assign_to_lists(oa, nlist)
    Qd = torch.from_numpy(Q).to("cuda:0")
    for topk, L in [(1, 2560), (10, 800), (3, 80)]:
        for tag, (gi, gd, gc) in (("split", grp.query_split(Qd, topk, L)), ("fused", grp.query(Qd, topk, L, "ivf"))):
            torch.cuda.synchronize()
            gi, gd, gc = gi.cpu().numpy(), gd.cpu().numpy(), gc.cpu().numpy()
            for bq, q in enumerate(Q):
                exp = O.query_ivf(O.dtable(q, cw, 16), codes, oc, offsets, ids, topk, L)
                n = int(gc[bq])
                assert_same_result(gi[bq, :n], gd[bq, :n], exp[0], exp[1], "%s M=%d nlist=%d k=%d L=%d" % (tag, M, nlist, topk, L))


@pytest.mark.parametrize("M,D", [(32, 128), (32, 96), (20, 40)])
def test_persistent_batch_kernel(M, D):
    """k_scan_persist32 (one resident CTA per SM, 11 scanning warps + 1 producer warp, double-buffered tables) must return
    exactly what the per-query kernel and the oracle return: fused coarse pass and given rankings (shards), every batch size
    around the grid size, topk up to 16, empty / flagged plans, huge tables (exact per-candidate path)."""
    import torch
    from rii_b200 import sharded
    Ks, N, nlist = 256, 70000, 120
    cw, codes, Q = synth(D, M, Ks, N, 24, seed=31 + D)
    e = engine(cw, codes)
    e.reconfigure(nlist, 1)
    centers = e.coarse_centers_array()
    offsets, ids = e.posting_lists_csr()
    for B in (1, 2, 3, 147, 149, 300, 613):
        Qb = np.ascontiguousarray(np.tile(Q, (B // len(Q) + 1, 1))[:B] + np.float32(0.001) * np.arange(B, dtype=np.float32)[:, None])
        for topk, L in [(1, 5000), (16, 700), (5, 40), (3, N)]:
            e.set_option("persist", 2)
            f = e.query_batch(Qb, topk, L=L, method="ivf")
            e.set_option("persist", 0)
            u = e.query_batch(Qb, topk, L=L, method="ivf")
            assert np.array_equal(f[2], u[2]) and np.array_equal(f[0], u[0]) and np.array_equal(bits(f[1]), bits(u[1])), (B, topk, L)
            for b in range(0, B, max(1, B // 5)):
                exp = O.query_ivf(O.dtable(Qb[b], cw, 16), codes, centers, offsets, ids, topk, L)
                n = int(f[2][b])
                assert_same_result(f[0][b][:n], f[1][b][:n], exp[0], exp[1], "persist B=%d k=%d L=%d b=%d" % (B, topk, L, b))
    # huge tables: queries ~1e19 away take the exact per-candidate path inside the same kernel
    Qh = np.ascontiguousarray(np.tile(Q, (13, 1))[:300])
    Qh[::7] *= np.float32(3e18)
    e.set_option("persist", 2)
    f = e.query_batch(Qh, 4, L=3000, method="ivf")
    e.set_option("persist", 0)
    u = e.query_batch(Qh, 4, L=3000, method="ivf")
    assert np.array_equal(f[2], u[2]) and np.array_equal(f[0], u[0]) and np.array_equal(bits(f[1]), bits(u[1]))
    # given rankings over two id-range shards (the plan uses glob_len / pre_len)
    if M == 32:
        grp = sharded.LocalShardGroup([engine(cw) for _ in range(2)])
        oc = grp.build(codes, nlist, 1)
        assert np.array_equal(oc, centers)
        Qd = torch.from_numpy(np.ascontiguousarray(np.tile(Q, (13, 1))[:300])).to("cuda:0")
        for eng in grp.engines:
            eng.e.set_option("persist", 2)
        for topk, L in [(1, 5000), (7, 333)]:
            gi, gd, gc = grp.query_split(Qd, topk, L)
            torch.cuda.synchronize()
            gi, gd, gc = gi.cpu().numpy(), gd.cpu().numpy(), gc.cpu().numpy()
            for b in range(0, 300, 37):
                exp = O.query_ivf(O.dtable(Qd[b].cpu().numpy(), cw, 16), codes, centers, offsets, ids, topk, L)
                n = int(gc[b])
                assert_same_result(gi[b, :n], gd[b, :n], exp[0], exp[1], "persist shards k=%d L=%d b=%d" % (topk, L, b))


def test_persistent_batch_kernel_exact_ties_across_lists():
    """topk = 1 on k_scan_persist32<K1> keeps one (distance, position) per warp; exact distance ties between candidates of
    DIFFERENT posting lists must still be decided by id (the reference ranks by distance only; we pin (distance, id)).
    Codes with 8 live sub-spaces over 4 codewords: 65536 distinct codes for 60000 rows -> duplicates and massive ties."""
    D, M, Ks, N, nlist = 64, 32, 4, 60000, 64
    cw, codes, Q = synth(D, M, Ks, N, 12, seed=77)
    codes[:, 8:] = 0
    e = engine(cw, codes)
    e.reconfigure(nlist, 1)
    centers = e.coarse_centers_array()
    offsets, ids = e.posting_lists_csr()
    B = 333
    Qb = np.ascontiguousarray(np.tile(Q, (B // len(Q) + 1, 1))[:B])
    for topk, L in [(1, 20000), (1, 900), (4, 20000)]:
        e.set_option("persist", 2)
        f = e.query_batch(Qb, topk, L=L, method="ivf")
        e.set_option("persist", 0)
        u = e.query_batch(Qb, topk, L=L, method="ivf")
        assert np.array_equal(f[2], u[2]) and np.array_equal(f[0], u[0]) and np.array_equal(bits(f[1]), bits(u[1])), (topk, L)
        for b in range(0, B, 29):
            exp = O.query_ivf(O.dtable(Qb[b], cw, 16), codes, centers, offsets, ids, topk, L)
            n = int(f[2][b])
            assert_same_result(f[0][b][:n], f[1][b][:n], exp[0], exp[1], "persist ties k=%d L=%d b=%d" % (topk, L, b))

def test_large_host_batch_is_pipelined_in_chunks():
    """rii_query_batch with >= 16384 host queries copies them in chunks on a second stream while the previous chunk is
    searched: the results must equal those of separate smaller calls, bit for bit (also across the chunk boundaries)."""
    cw, codes, Q = synth(128, 32, 256, 60000, 50, seed=5)
    e = engine(cw, codes)
    e.reconfigure(100, 1)
    B = 8192 * 2 + 1234
    Qb = np.ascontiguousarray(np.tile(Q, (B // len(Q) + 1, 1))[:B] + np.float32(1e-4) * (np.arange(B, dtype=np.float32) % 977)[:, None])
    for topk, L, method in [(1, 3000, "ivf"), (3, 3000, "ivf")]:
        big = e.query_batch(Qb, topk, L=L, method=method)
        for s0 in (0, 8192 - 100, 16384 - 50, B - 300):
            small = e.query_batch(np.ascontiguousarray(Qb[s0:s0 + 300]), topk, L=L, method=method)
            n = min(300, B - s0)
            assert np.array_equal(big[0][s0:s0 + n], small[0][:n]) and np.array_equal(bits(big[1][s0:s0 + n]), bits(small[1][:n]))
            assert np.array_equal(big[2][s0:s0 + n], small[2][:n])


def test_randomised_shapes_against_the_oracle():
    """40 random (M, Ks, N, nlist, batch, topk, L) draws through the automatic dispatch: every combination of engine
    (persistent / fused multi-CTA with the merge by the last CTA / k_merge / natural layout / general path), batch size and
    top-k mode (topk = 1 instantiations, warp lists) must agree with the oracle bit for bit."""
    rng = np.random.default_rng(2026)
    for trial in range(40):
        M = int(rng.choice([8, 16, 20, 32, 32, 32, 48, 64]))
        Ds = int(rng.choice([1, 2, 4]))
        Ks = int(rng.choice([16, 64, 256, 256]))
        N = int(rng.choice([3000, 40000, 90000]))
        nlist = int(rng.choice([7, 60, 300]))
        B = int(rng.choice([1, 2, 5, 40, 150, 310]))
        topk = int(rng.choice([1, 1, 1, 2, 9, 33]))
        L = int(rng.choice([topk, 500, 5000, N]))
        L = max(topk, min(L, N))
        cw, codes, Q = synth(M * Ds, M, Ks, N, 16, seed=1000 + trial)
        if trial % 5 == 0:
            codes[:, 3:] = 0  # exact ties
        e = engine(cw, codes)
        e.reconfigure(nlist, 1)
        centers = e.coarse_centers_array()
        offsets, ids = e.posting_lists_csr()
        Qb = np.ascontiguousarray(np.tile(Q, (B // len(Q) + 1, 1))[:B])
        what = "trial %d: M=%d Ds=%d Ks=%d N=%d nlist=%d B=%d topk=%d L=%d" % (trial, M, Ds, Ks, N, nlist, B, topk, L)
        gi, gd, gc = e.query_batch(Qb, topk, L=L, method="ivf")
        li, ld, lc = e.query_batch(Qb, topk, method="linear")
        for b in sorted(set([0, B // 2, B - 1])):
            T = O.dtable(Qb[b], cw, 16)
            exp = O.query_ivf(T, codes, centers, offsets, ids, topk, L)
            n = int(gc[b])
            assert_same_result(gi[b][:n], gd[b][:n], exp[0], exp[1], what + " ivf b=%d" % b)
            exp = O.query_linear(T, codes, topk)
            assert int(lc[b]) == topk
            assert_same_result(li[b], ld[b], exp[0], exp[1], what + " linear b=%d" % b)


def test_opq_rotation_on_the_device_and_small_call_path():
    """(1) rii/rii.py:305-306: the OPQ rotation of the query folded into the engine (k_rotate, fp32 FMA chain) returns the
    ids of the host rotation and distances within the 1e-5 relative contract; it survives pickling.  (2) single calls
    (zero-copy through mapped pinned memory) == the staged-copy path."""
    import copy
    from rii_b200 import Rii, pq
    rng = np.random.default_rng(9)
    D, M, N = 64, 16, 20000
    X = rng.random((N, D), dtype=np.float32)
    codec = pq.OPQ(M=M, Ks=256, verbose=False).fit(X[:2000], pq_iter=3, rotation_iter=2, seed=123)
    e_host = Rii(codec).add_configure(X, nlist=50, iter=2)
    e_dev = Rii(codec, rotate_on_device=True).add_configure(X, nlist=50, iter=2)
    e_dev2 = copy.deepcopy(e_dev)
    Q = rng.random((20, D), dtype=np.float32)
    for q in Q:
        for method in ("linear", "ivf"):
            ih, dh = e_host.query(q, topk=5, L=2000, method=method)
            for e in (e_dev, e_dev2):
                i2, d2 = e.query(q, topk=5, L=2000, method=method)
                assert np.allclose(d2, dh, rtol=1e-5, atol=0), (method, d2, dh)   # tolerance: north_star's 1e-5 relative
                assert np.array_equal(i2, ih) or np.allclose(np.sort(d2), np.sort(dh), rtol=1e-5)
    bi, bd, bc = e_dev.query_batch(Q, topk=5, L=2000, method="ivf")
    hi, hd, hc = e_host.query_batch(Q, topk=5, L=2000, method="ivf")
    assert np.allclose(bd, hd, rtol=1e-5) and (bi == hi).mean() > 0.95
    # zero-copy small calls == staged copies, bit for bit
    imp = e_host.impl_cpp
    qr = codec.rotate(Q)
    for q in qr[:8]:
        imp.set_option("zero_copy", 1)
        a = (imp.query_linear(q, 7, EMPTY), imp.query_ivf(q, 7, EMPTY, 1500))
        imp.set_option("zero_copy", 0)
        b = (imp.query_linear(q, 7, EMPTY), imp.query_ivf(q, 7, EMPTY, 1500))
        assert a == b
    imp.set_option("zero_copy", 1)
