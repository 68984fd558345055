"""Shared helpers of the test-suite (oracle access, golden fixtures, canonical comparisons)."""
import glob
import os

import numpy as np

from oracle import oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files(build="strict_v4"):
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.%s.npz" % build)))


def load_golden(path):
    z = np.load(path)
    g = {k: z[k] for k in z.files}
    g["name"] = os.path.basename(path)
    g["nlist"], g["iter"], g["variant"] = int(g["nlist"]), int(g["iter"]), int(g["variant"])
    return g


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_same_result(ids, dists, exp_ids, exp_dists, what=""):
    """ids bit-exact, distances bit-exact (fp32)."""
    ids, exp_ids = np.asarray(ids, np.int64), np.asarray(exp_ids, np.int64)
    assert ids.shape == exp_ids.shape, "%s: count %s vs %s" % (what, ids.shape, exp_ids.shape)
    assert np.array_equal(ids, exp_ids), "%s: ids differ\n got %s\n exp %s" % (what, ids[:10], exp_ids[:10])
    assert np.array_equal(bits(dists), bits(exp_dists)), "%s: distances differ" % what


def canonical_topk(ids, dists, k):
    ids = np.asarray(ids, np.int64)
    dists = np.asarray(dists, np.float32)
    o = np.lexsort((ids, dists))[:k]
    return ids[o], dists[o]


def synth(D, M, Ks, N, nq, seed):
    """Codewords / codes / queries the way the reference's tests and README draw them (uniform [0,1))."""
    rng = np.random.default_rng(seed)
    cw = rng.random((M, Ks, D // M), dtype=np.float32)
    codes = rng.integers(0, Ks, (N, M), dtype=np.uint8)
    Q = rng.random((nq, D), dtype=np.float32)
    return cw, codes, Q


__all__ = ["O", "golden_files", "load_golden", "bits", "assert_same_result", "canonical_topk", "synth"]
