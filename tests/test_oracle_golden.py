"""CPU: the oracle restatement (oracle/rii_oracle.cpp) against the golden vectors recorded from the
unmodified reference (tests/golden/*.npz, generator tests/golden/make_golden.py).  Bit-exact."""
import numpy as np
import pytest

from _util import O, assert_same_result, bits, canonical_topk, golden_files, load_golden

FILES = golden_files("strict_v4") + golden_files("strict_v3")


def test_fixtures_present():
    assert len(FILES) >= 7


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1])
def test_dtable_and_adist(path):
    g = load_golden(path)
    for i, q in enumerate(g["Q"]):
        T = O.dtable(q, g["cw"], g["variant"])
        d = O.adist_all(T, g["codes"])
        assert np.array_equal(bits(d), bits(g["all_dists"][i])), "%s query %d" % (g["name"], i)


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1])
def test_query_linear(path):
    g = load_golden(path)
    for j, (i, topk) in enumerate(g["lin_meta"]):
        T = O.dtable(g["Q"][i], g["cw"], g["variant"])
        ids, d = O.query_linear(T, g["codes"], int(topk))
        # canonical expectation from the recorded full distance vector
        eid, ed = canonical_topk(np.arange(len(g["codes"])), g["all_dists"][i], int(topk))
        assert_same_result(ids, d, eid, ed, "linear")
        # the reference's own top-k: same distances (tie order among equal distances is unspecified there)
        assert np.array_equal(bits(np.sort(g["lin_%d_d" % j])), bits(d))
        sids, sd = O.query_linear(T, g["codes"], int(topk), g["tids"])
        eid, ed = canonical_topk(g["tids"], g["all_dists"][i][g["tids"]], int(topk))
        assert_same_result(sids, sd, eid, ed, "linear subset")
        assert np.array_equal(bits(np.sort(g["lin_%d_sd" % j])), bits(sd))


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1])
def test_query_ivf(path):
    g = load_golden(path)
    assert len(g["ivf_meta"]) > 0
    for j, (i, topk, L, sub, ncand) in enumerate(g["ivf_meta"]):
        T = O.dtable(g["Q"][i], g["cw"], g["variant"])
        ids, d, nc = O.query_ivf(T, g["codes"], g["centers"], g["offsets"], g["ids"], int(topk), int(L),
                                 g["tids"] if sub else None, return_ncand=True)
        assert nc == ncand
        assert_same_result(ids, d, g["ivf_%d_ids" % j], g["ivf_%d_d" % j], "ivf case %d" % j)
        assert np.array_equal(bits(np.sort(g["ivf_%d_raw_d" % j])), bits(d))


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1])
def test_reconfigure(path):
    g = load_golden(path)
    centers, assign = O.reconfigure(g["cw"], g["codes"], g["nlist"], g["iter"])
    assert np.array_equal(centers, g["centers"])
    offsets, ids = O.assign_to_lists(assign, g["nlist"])
    assert np.array_equal(offsets, g["offsets"])
    assert np.array_equal(ids, g["ids"])


def test_reference_pickle_stream_loads_in_the_unmodified_reference():
    """SURVEY 8f rank 3 / src/main.cpp:35-54: the 5-tuple state our index exports (RiiCpp.to_reference_state /
    dumps_reference) must unpickle into the reference's own main.RiiCpp and answer queries there."""
    import os
    import pickle
    import subprocess
    import sys
    import tempfile
    from oracle import ref as R
    from rii_b200.main import reference_pickle_bytes
    if not R.available("strict"):
        import pytest
        pytest.skip("oracle/_ref is not built here")
    g = load_golden(golden_files("strict_v4")[0] if golden_files("strict_v4") else golden_files("strict_v3")[0])
    centers, codes = g["centers"], g["codes"]
    offsets, ids = g["offsets"], g["ids"]
    pl = [ids[offsets[i]:offsets[i + 1]].tolist() for i in range(len(offsets) - 1)]
    blob = reference_pickle_bytes(g["cw"].tolist(), False, centers.tolist(), codes.reshape(-1).tolist(), pl)
    with tempfile.TemporaryDirectory() as td:
        f = os.path.join(td, "idx.pkl")
        open(f, "wb").write(blob)
        so_dir = R.variant_dir("strict")
        code = ("import sys, pickle, numpy as np; sys.path.insert(0, %r); import main; e = pickle.load(open(%r, 'rb'));"
                "assert type(e).__name__ == 'RiiCpp' and e.N == %d and e.nlist == %d;"
                "q = np.load(%r); r = e.query_linear(q, 3, np.array([], np.int64)); print(r[0])"
                % (so_dir, f, codes.shape[0], len(pl), os.path.join(td, "q.npy")))
        np.save(os.path.join(td, "q.npy"), np.ascontiguousarray(g["Q"][0], np.float32))
        out = subprocess.check_output([sys.executable, "-c", code], stderr=subprocess.STDOUT).decode()
    T = O.dtable(g["Q"][0], g["cw"], g["variant"])
    exp = O.query_linear(T, codes, 3)
    assert str(exp[0].tolist()) in out, out
