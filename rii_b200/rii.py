"""`rii.Rii` public API (rii/rii.py:6-400 of the reference) hosted on the B200 ADC path.

Same properties, methods, argument meaning, assertions and return dtypes; the only engine underneath is
`rii_b200.main.RiiCpp` (CUDA, via the C ABI).  Additions are opt-in: `query_batch`, `device=`.
"""
import copy

import numpy as np

from . import main
from . import pq as _pq
from .cost_model import CostModel, Threshold


class Rii(object):
    """Reconfigurable inverted index over PQ codes (IVFADC with subset search and reconfiguration).

    Args:
        fine_quantizer: a trained PQ / OPQ instance (`rii_b200.pq` or `nanopq`); rii/rii.py:32-38.
        device (int): CUDA device ordinal of the index.
    """

    def __init__(self, fine_quantizer, device=0, rotate_on_device=False):
        assert _pq.is_quantizer(fine_quantizer)
        assert fine_quantizer.codewords is not None, "Please fit the PQ/OPQ instance first"
        assert fine_quantizer.Ks <= 256, "Ks must be less than 256 so that each code must be uint8"
        self.fine_quantizer = copy.deepcopy(fine_quantizer)
        self.impl_cpp = main.RiiCpp(fine_quantizer.codewords, fine_quantizer.verbose, device=device)
        self.threshold = None
        # OPQ (rii/rii.py:305-306 rotates every query on the host): optionally folded into the engine, ahead of the table build
        self._device_rotation = bool(rotate_on_device) and _pq.is_opq(fine_quantizer)
        if self._device_rotation:
            self.impl_cpp.set_rotation(fine_quantizer.R)

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._device_rotation = bool(state.get("_device_rotation", False))
        if self._device_rotation:  # the engine's pickle state is the reference's (src/main.cpp:35-54): the rotation is ours to restore
            self.impl_cpp.set_rotation(self.fine_quantizer.R)

    # ---- properties, rii/rii.py:40-121 ------------------------------------------------------
    @property
    def M(self):
        return self.fine_quantizer.M

    @property
    def Ks(self):
        return self.fine_quantizer.Ks

    @property
    def N(self):
        return self.impl_cpp.N

    @property
    def nlist(self):
        return self.impl_cpp.nlist

    @property
    def codewords(self):
        return self.fine_quantizer.codewords

    @property
    def coarse_centers(self):
        if self.nlist == 0:
            return None
        return self.impl_cpp.coarse_centers_array().astype(self.fine_quantizer.code_dtype)

    @property
    def codes(self):
        if self.N == 0:
            return None
        return self.impl_cpp.codes_array().astype(self.fine_quantizer.code_dtype).reshape(self.N, self.M)

    @property
    def posting_lists(self):
        return self.impl_cpp.posting_lists

    @property
    def verbose(self):
        return self.impl_cpp.verbose

    @verbose.setter
    def verbose(self, v):
        self.fine_quantizer.verbose = v
        self.impl_cpp.verbose = v

    @property
    def L0(self):
        if self.nlist == 0:
            return None
        return int(np.round(self.N / self.nlist))

    # ---- build, rii/rii.py:123-233 ----------------------------------------------------------
    def reconfigure(self, nlist=None, iter=5):
        if nlist is None:
            nlist = int(np.sqrt(self.N))
        assert 0 < nlist
        self.impl_cpp.reconfigure(nlist, iter)
        self.threshold = estimate_best_threshold_function(e=self)

    def add(self, vecs, update_posting_lists="auto", gpu_encode=False):
        """rii/rii.py:152-186.  gpu_encode=True (opt-in, not in the reference) encodes on the GPU (rii_encode) instead
        of fine_quantizer.encode; same nearest-codeword rule, rounding may differ from nanopq on exact ties."""
        assert vecs.ndim == 2
        assert vecs.dtype == np.float32
        if gpu_encode:
            v = self.fine_quantizer.rotate(vecs) if _pq.is_opq(self.fine_quantizer) else vecs
            codes = self.impl_cpp.encode(v)
        else:
            codes = self.fine_quantizer.encode(vecs)
        self.impl_cpp.add_codes(codes, self._resolve_update_posting_lists_flag(update_posting_lists))

    def add_configure(self, vecs, nlist=None, iter=5):
        self.add(vecs=vecs, update_posting_lists=False)
        self.reconfigure(nlist=nlist, iter=iter)
        return self

    def merge(self, engine, update_posting_lists="auto"):
        assert isinstance(engine, Rii)
        assert self.fine_quantizer == engine.fine_quantizer, \
            "Two engines to be merged must have the same fine quantizer"
        if engine.N != 0:
            self.impl_cpp.add_codes(engine.codes, self._resolve_update_posting_lists_flag(update_posting_lists))
        if self.verbose:
            print("The number of codes: {}".format(self.N))

    # ---- search, rii/rii.py:235-320 ---------------------------------------------------------
    def _prepare(self, topk, L, target_ids, sort_target_ids):
        assert 0 < self.N
        assert 0 < self.nlist
        if topk is None:
            topk = self.N
        assert 1 <= topk <= self.N
        if L is None:
            L = self._multiple_of_L0_covering_topk(topk=topk)
        assert topk <= L <= self.N, \
            "Parameters are weird. Make sure topk<=L<=N:  topk={}, L={}, N={}".format(topk, L, self.N)
        if target_ids is None:
            tids = np.array([], dtype=np.int64)
            len_target_ids = self.N
        else:
            assert isinstance(target_ids, np.ndarray)
            assert target_ids.dtype == np.int64
            assert target_ids.ndim == 1
            tids = np.sort(target_ids) if sort_target_ids else np.ascontiguousarray(target_ids)
            len_target_ids = len(tids)
        assert topk <= len_target_ids <= self.N, \
            "Parameters are weird. Make sure topk<=len(target_ids)<=N:  " \
            "topk={}, len(target_ids)={}, N={}".format(topk, len_target_ids, self.N)
        return topk, L, tids, len_target_ids

    def query(self, q, topk=1, L=None, target_ids=None, sort_target_ids=True, method="auto"):
        """ids (topk,) int64 and distances (topk,) float64 of the nearest PQ codes; rii/rii.py:235-320."""
        assert method in ["auto", "linear", "ivf"]
        topk, L, tids, len_target_ids = self._prepare(topk, L, target_ids, sort_target_ids)
        q_ = self.fine_quantizer.rotate(q) if _pq.is_opq(self.fine_quantizer) and not self._device_rotation else q
        if method == "auto":
            method = "linear" if self._use_linear(len_target_ids, L, subset=target_ids is not None) else "ivf"
        if method == "linear":
            ids, dists = self.impl_cpp.query_linear(q_, topk, tids)
        else:
            ids, dists = self.impl_cpp.query_ivf(q_, topk, tids, L)
        return np.array(ids, np.int64), np.array(dists)

    def query_batch(self, Q, topk=1, L=None, target_ids=None, sort_target_ids=True, method="ivf"):
        """Batch form of :func:`query` (not in the reference): Q (B, D) float32 ->
        ids (B, topk) int64 (-1 padded), dists (B, topk) float64 (inf padded), counts (B,) int32."""
        assert method in ["auto", "linear", "ivf"]
        assert Q.ndim == 2 and Q.dtype == np.float32
        topk, L, tids, _ = self._prepare(topk, L, target_ids, sort_target_ids)
        Q_ = self.fine_quantizer.rotate(Q) if _pq.is_opq(self.fine_quantizer) and not self._device_rotation else Q
        if method == "auto":
            method = "linear" if self._use_linear(len(tids) if target_ids is not None else self.N, L,
                                                   subset=target_ids is not None, batch=Q.shape[0]) else "ivf"
        ids, dists, counts = self.impl_cpp.query_batch(np.ascontiguousarray(Q_, np.float32), topk, tids, L, method)
        return ids, dists.astype(np.float64), counts

    def clear(self):
        self.threshold = None
        self.impl_cpp.clear()

    def print_params(self):
        print("verbose:", self.verbose)
        print("M:", self.M)
        print("Ks:", self.Ks)
        print("fine_quantizer:", self.fine_quantizer)
        print("N:", self.N)
        print("nlist:", self.nlist)
        print("L0:", self.L0)
        print("cordwords.shape:", self.codewords.shape)
        print("coarse_centers.shape:", None if self.nlist == 0 else self.coarse_centers.shape)
        print("codes.shape:", None if self.codes is None else self.codes.shape)
        lens = [len(p) for p in self.posting_lists]
        print("[len(poslist) for poslist in posting_lists]: [" + "".join(str(v) + ", " for v in lens[:11]) +
              (" ..." if len(lens) > 11 else "") + "]")
        for topk in 1, 10, 100:
            L = "None" if self.nlist == 0 else self._multiple_of_L0_covering_topk(topk)
            print("_multiple_of_L0_covering_topk(topk={}): {}".format(topk, L))
        print("threshold function thre_{|S|}=f(L):", self.threshold)
        for S in [10 ** (2 + n) for n in range(5)]:
            use_linear = None if self.threshold is None else self._use_linear(S, self.L0, subset=True)
            print("_use_linear({S}, L={L0}): {use_linear}".format(S=S, L0=self.L0, use_linear=use_linear))

    # ---- helpers, rii/rii.py:374-400 --------------------------------------------------------
    def _multiple_of_L0_covering_topk(self, topk):
        return min((topk // self.L0 + 1) * self.L0, self.N)

    def _use_linear(self, len_target_ids, L, subset=None, batch=1):
        """rii/rii.py:383-392 (`len_target_ids <= threshold(L)`), decided by the cost model: with target_ids the
        comparison is |S| against the crossover thre(L); without, N candidates against nlist + L."""
        if subset is None:
            subset = len_target_ids != self.N
        if not subset:
            return bool(self.threshold.model.use_linear(self.N, L, False, batch))
        return bool(len_target_ids <= self.threshold(L, batch))

    def _resolve_update_posting_lists_flag(self, flag):
        assert flag in ["auto", True, False]
        if flag == "auto":
            return 0 < self.nlist
        return flag


def estimate_best_threshold_function(e, queries=None):
    """thre_{|S|} = f(L): the |S| at which a linear scan over target_ids and the inverted-index search cost the same
    (rii/rii.py:403-486).  The reference measures wall-clock times of both methods and fits a line; here the crossover
    comes from a bytes-and-launches cost model (rii_b200/cost_model.py), so it is deterministic and costs no queries.
    `queries` is accepted for signature compatibility and ignored."""
    t = Threshold(CostModel(e.N, e.nlist, e.M))
    if e.verbose:
        print("===== Threshold selection ====")
        Ls = [k * e._multiple_of_L0_covering_topk(k) for k in [1, 2, 4, 8, 16] if k * e._multiple_of_L0_covering_topk(k) <= e.N]
        print("L:", Ls)
        print("threshold:", [t(L) for L in Ls])
    return t
