"""Host-side product quantizers with the `nanopq.PQ` / `nanopq.OPQ` surface that `rii.Rii` relies on.

nanopq is a third-party dependency of the reference (requirements.txt:2) that is neither vendored under
/root/reference nor installed here; it sits *outside* the ADC hot path (codebook training, encode, decode,
OPQ rotation: call sites rii/rii.py:33-37,150,185,225,305-306).  These classes give `Rii` something to be
constructed from.  Contract taken from those call sites: `.fit(vecs)`, `.encode`,
`.decode`, `.codewords` (M, Ks, Ds) float32, `.M`, `.Ks`, `.Ds`, `.verbose`, `.code_dtype`, `__eq__`, and
`OPQ.rotate`.  If the real nanopq is installed, `rii_b200.Rii` accepts its instances just the same.
Encoding parity with nanopq is NOT pinned (nothing in the reference pins nanopq's numeric output).
"""
import numpy as np
from scipy.cluster.vq import kmeans2, vq


class PQ(object):
    def __init__(self, M, Ks=256, metric="l2", verbose=True):
        assert 0 < Ks <= 2 ** 32
        assert metric == "l2", "only the l2 metric is on Rii's path"
        self.M, self.Ks, self.metric, self.verbose = M, Ks, metric, verbose
        self.code_dtype = np.uint8 if Ks <= 2 ** 8 else (np.uint16 if Ks <= 2 ** 16 else np.uint32)
        self.codewords = None
        self.Ds = None
        if verbose:
            print("M: {}, Ks: {}, metric : {}, code_dtype: {}".format(M, Ks, metric, self.code_dtype))

    def __eq__(self, other):
        if not isinstance(other, PQ):
            return False
        same = (self.M, self.Ks, self.metric, self.verbose, self.code_dtype, self.Ds) == \
               (other.M, other.Ks, other.metric, other.verbose, other.code_dtype, other.Ds)
        if not same:
            return False
        if self.codewords is None or other.codewords is None:
            return self.codewords is other.codewords
        return np.array_equal(self.codewords, other.codewords)

    __hash__ = None

    def fit(self, vecs, iter=20, seed=123, minit="points"):
        assert vecs.dtype == np.float32 and vecs.ndim == 2
        N, D = vecs.shape
        assert self.Ks < N, "the number of training vectors should be more than Ks"
        assert D % self.M == 0, "input dimension must be dividable by M"
        self.Ds = D // self.M
        if self.verbose:
            print("iter: {}, seed: {}".format(iter, seed))
        self.codewords = np.zeros((self.M, self.Ks, self.Ds), dtype=np.float32)
        for m in range(self.M):
            if self.verbose:
                print("Training the subspace: {} / {}".format(m, self.M))
            sub = vecs[:, m * self.Ds:(m + 1) * self.Ds]
            self.codewords[m], _ = kmeans2(sub, self.Ks, iter=iter, minit=minit, seed=seed)
        return self

    def encode(self, vecs):
        assert vecs.dtype == np.float32 and vecs.ndim == 2
        N, D = vecs.shape
        assert D == self.Ds * self.M, "input dimension must be Ds * M"
        codes = np.empty((N, self.M), dtype=self.code_dtype)
        for m in range(self.M):
            if self.verbose:
                print("Encoding the subspace: {} / {}".format(m, self.M))
            codes[:, m], _ = vq(vecs[:, m * self.Ds:(m + 1) * self.Ds], self.codewords[m])
        return codes

    def decode(self, codes):
        assert codes.ndim == 2 and codes.shape[1] == self.M
        vecs = np.empty((codes.shape[0], self.Ds * self.M), dtype=np.float32)
        for m in range(self.M):
            vecs[:, m * self.Ds:(m + 1) * self.Ds] = self.codewords[m][codes[:, m], :]
        return vecs


class OPQ(object):
    """PQ preceded by a learned orthogonal rotation R (non-parametric OPQ: alternate PQ training and an
    orthogonal Procrustes solve)."""

    def __init__(self, M, Ks=256, metric="l2", verbose=True):
        self.pq = PQ(M, Ks, metric=metric, verbose=verbose)
        self.R = None

    def __eq__(self, other):
        if not isinstance(other, OPQ):
            return False
        if self.pq != other.pq:
            return False
        if self.R is None or other.R is None:
            return self.R is other.R
        return np.array_equal(self.R, other.R)

    __hash__ = None

    M = property(lambda self: self.pq.M)
    Ks = property(lambda self: self.pq.Ks)
    Ds = property(lambda self: self.pq.Ds)
    code_dtype = property(lambda self: self.pq.code_dtype)
    codewords = property(lambda self: self.pq.codewords)

    @property
    def verbose(self):
        return self.pq.verbose

    @verbose.setter
    def verbose(self, v):
        self.pq.verbose = v

    def fit(self, vecs, pq_iter=20, rotation_iter=10, seed=123, minit="points"):
        assert vecs.dtype == np.float32 and vecs.ndim == 2
        D = vecs.shape[1]
        R = np.eye(D, dtype=np.float32)
        verbose = self.pq.verbose
        for i in range(rotation_iter):
            if verbose:
                print("OPQ rotation training: {} / {}".format(i, rotation_iter))
            X = vecs @ R
            last = i == rotation_iter - 1
            self.pq.verbose = verbose and last
            pq_tmp = PQ(self.M, self.Ks, verbose=False).fit(X, iter=pq_iter if last else 1, seed=seed, minit=minit)
            if last:
                self.pq.codewords, self.pq.Ds = pq_tmp.codewords, pq_tmp.Ds
                break
            X_ = pq_tmp.decode(pq_tmp.encode(X))
            U, _, Vt = np.linalg.svd(vecs.T @ X_)
            R = (U @ Vt).astype(np.float32)
        self.pq.verbose = verbose
        self.R = R
        return self

    def rotate(self, vecs):
        assert vecs.ndim in (1, 2)
        if vecs.ndim == 2:
            return (vecs @ self.R).astype(np.float32)
        return (vecs.reshape(1, -1) @ self.R).reshape(-1).astype(np.float32)

    def encode(self, vecs):
        return self.pq.encode(self.rotate(vecs))

    def decode(self, codes):
        return (self.pq.decode(codes) @ self.R.T).astype(np.float32)


def is_quantizer(obj):
    """isinstance check of rii/rii.py:33, also accepting the real nanopq classes when installed."""
    if isinstance(obj, (PQ, OPQ)):
        return True
    try:
        import nanopq  # noqa
        return isinstance(obj, (nanopq.PQ, nanopq.OPQ))
    except ImportError:
        return False


def is_opq(obj):
    if isinstance(obj, OPQ):
        return True
    try:
        import nanopq  # noqa
        return isinstance(obj, nanopq.OPQ)
    except ImportError:
        return False
