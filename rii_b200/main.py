"""`main.RiiCpp` work-alike over the C ABI: same constructor, methods, attributes, return types and dtype
strictness as the reference's pybind11 class (src/main.cpp:12-54), so that rii/rii.py-style host code and the
reference's own tests (tests/test_rii.py) run unchanged against the B200 path."""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import check

__version__ = "0.2.12"


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def _strict(a, dtype, ndim, name):
    """py::arg(name).noconvert() (src/main.cpp:18-26): wrong dtype / layout raises TypeError."""
    if not isinstance(a, np.ndarray) or a.dtype != dtype or a.ndim != ndim or not a.flags.c_contiguous:
        raise TypeError("%s must be a C-contiguous numpy.ndarray[%s] with ndim=%d (no implicit conversion)"
                        % (name, np.dtype(dtype).name, ndim))
    return a


class RiiCpp(object):
    def __init__(self, codewords=None, verbose=False, device=0, l2_variant=0):
        self._h = None
        self._outs = {}
        self._lib = _capi.lib()
        self._codewords = None
        self._device = device
        self._l2_variant = l2_variant
        if codewords is not None:
            self._create(codewords, verbose)

    def _create(self, codewords, verbose):
        cw = np.ascontiguousarray(codewords, dtype=np.float32)  # src/main.cpp:14 allows implicit conversion
        if cw.ndim != 3:
            raise ValueError("codewords must have ndim=3: (M, Ks, Ds)")
        self._codewords = cw
        M, Ks, Ds = cw.shape
        h = C.c_void_p()
        check(_capi.lib().rii_create(_ptr(cw, C.c_float), M, Ks, Ds, int(bool(verbose)), self._device,
                                     self._l2_variant, C.byref(h)))
        self._h = h
        self.M, self.Ks, self.Ds = M, Ks, Ds

    def __del__(self):
        try:
            if self._h is not None:
                _capi.lib().rii_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- src/main.cpp:15-28 -------------------------------------------------------------------
    def reconfigure(self, nlist, iter):
        check(_capi.lib().rii_reconfigure(self._h, int(nlist), int(iter)))

    def add_codes(self, codes, update_flag):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        if codes.ndim != 2 or codes.shape[1] != self.M:
            raise ValueError("codes must have shape (N, M)")
        check(_capi.lib().rii_add_codes(self._h, _ptr(codes, C.c_uint8), codes.shape[0], int(bool(update_flag))))

    def _out(self, topk):
        """Result buffers of a single-query call, kept per topk with their ctypes pointers: a call is a few microseconds of
        device work, so the wrapper must not spend more than that on allocations and pointer conversions."""
        o = self._outs.get(topk)
        if o is None:
            ids, dists = np.empty(topk, np.int64), np.empty(topk, np.float32)
            o = self._outs[topk] = (ids, dists, ids.ctypes.data, dists.ctypes.data)
        return o

    def _single(self, query, topk, target_ids, L, ivf):
        q = _strict(query, np.float32, 1, "query")
        t = _strict(target_ids, np.int64, 1, "target_ids")
        if q.shape[0] != self.M * self.Ds:
            raise ValueError("query must have M * Ds = %d elements" % (self.M * self.Ds))
        topk = int(topk)
        ids, dists, pi, pd = self._out(topk)
        pq, pt = q.ctypes.data, (t.ctypes.data if t.size else None)
        lib = self._lib
        if ivf:
            n = lib.rii_query_ivf(self._h, pq, topk, pt, t.size, int(L), pi, pd)
        else:
            n = lib.rii_query_linear(self._h, pq, topk, pt, t.size, pi, pd)
        if n < 0:
            check(n)
        return ids[:n].tolist(), dists[:n].tolist()

    def query_linear(self, query, topk, target_ids):
        return self._single(query, topk, target_ids, 0, False)

    def query_ivf(self, query, topk, target_ids, L):
        return self._single(query, topk, target_ids, L, True)

    def clear(self):
        check(_capi.lib().rii_clear(self._h))

    # ---- batch entry (new; include/rii_b200.h rii_query_batch) -------------------------------
    def query_batch(self, queries, topk, target_ids=None, L=0, method="linear"):
        """queries float32 (B, D) -> (ids int64 (B, topk), dists float32 (B, topk), counts int32 (B))."""
        Q = _strict(queries, np.float32, 2, "queries")
        t = np.empty(0, np.int64) if target_ids is None else _strict(target_ids, np.int64, 1, "target_ids")
        B = Q.shape[0]
        ids = np.full((B, int(topk)), -1, np.int64)
        dists = np.full((B, int(topk)), np.inf, np.float32)
        counts = np.zeros(B, np.int32)
        m = {"linear": 0, "ivf": 1}[method]
        check(_capi.lib().rii_query_batch(self._h, _ptr(Q, C.c_float), B, int(topk), _ptr(t, C.c_int64), t.size,
                                          int(L), m, _ptr(ids, C.c_int64), _ptr(dists, C.c_float),
                                          _ptr(counts, C.c_int32)))
        return ids, dists, counts

    # ---- attributes, src/main.cpp:29-34 -----------------------------------------------------
    @property
    def verbose(self):
        return bool(_capi.lib().rii_get_verbose(self._h))

    @verbose.setter
    def verbose(self, v):
        check(_capi.lib().rii_set_verbose(self._h, int(bool(v))))

    @property
    def N(self):
        return int(_capi.lib().rii_get_N(self._h)) if self._h is not None else 0

    @property
    def nlist(self):
        return int(_capi.lib().rii_get_nlist(self._h)) if self._h is not None else 0

    # numpy views of the state (the list-returning properties below copy-convert like the reference)
    def codes_array(self):
        out = np.empty((self.N, self.M), np.uint8)
        if self.N:
            check(_capi.lib().rii_copy_codes(self._h, _ptr(out, C.c_uint8)))
        return out

    def coarse_centers_array(self):
        out = np.empty((self.nlist, self.M), np.uint8)
        if self.nlist:
            check(_capi.lib().rii_copy_coarse_centers(self._h, _ptr(out, C.c_uint8)))
        return out

    def posting_lists_csr(self):
        offsets = np.zeros(self.nlist + 1, np.int64)
        ids = np.empty(self.N, np.int32)
        check(_capi.lib().rii_copy_posting_lists(self._h, _ptr(offsets, C.c_int64), _ptr(ids, C.c_int32)))
        return offsets, ids[: offsets[-1]]

    @property
    def coarse_centers(self):
        return self.coarse_centers_array().tolist()

    @property
    def flattened_codes(self):
        return self.codes_array().reshape(-1).tolist()

    @property
    def posting_lists(self):
        offsets, ids = self.posting_lists_csr()
        return [ids[offsets[i]:offsets[i + 1]].tolist() for i in range(self.nlist)]

    # ---- pickle, src/main.cpp:35-54: the same 5-tuple, with arrays instead of nested lists ---
    def __getstate__(self):
        offsets, ids = self.posting_lists_csr()
        return (self._codewords, self.verbose, self.coarse_centers_array(), self.codes_array(), (offsets, ids),
                self._device, self._l2_variant)

    def __setstate__(self, t):
        if len(t) not in (5, 7):
            raise RuntimeError("Invalid state when reading pickled item")
        self._h = None
        self._outs = {}
        self._lib = _capi.lib()
        self._device, self._l2_variant = (t[5], t[6]) if len(t) == 7 else (0, 0)
        self._create(np.asarray(t[0], np.float32), bool(t[1]))
        centers = np.ascontiguousarray(t[2], np.uint8).reshape(-1, self.M)
        codes = np.ascontiguousarray(t[3], np.uint8).reshape(-1, self.M)
        pl = t[4]
        if isinstance(pl, tuple):
            offsets, ids = np.ascontiguousarray(pl[0], np.int64), np.ascontiguousarray(pl[1], np.int32)
        else:  # the reference's list-of-lists form
            offsets = np.zeros(len(pl) + 1, np.int64)
            offsets[1:] = np.cumsum([len(p) for p in pl])
            ids = np.ascontiguousarray(np.concatenate([np.asarray(p, np.int32) for p in pl]) if offsets[-1] else
                                       np.zeros(0, np.int32), np.int32)
        check(_capi.lib().rii_set_state(self._h, _ptr(centers, C.c_uint8), centers.shape[0], _ptr(codes, C.c_uint8),
                                        codes.shape[0], _ptr(offsets, C.c_int64), _ptr(ids, C.c_int32)))

    # ---- exchange with the reference (src/main.cpp:35-54) ------------------------------------------
    def to_reference_state(self):
        """The reference's own pickle state, exactly: (codewords, verbose, coarse_centers, flattened_codes,
        posting_lists) as nested Python lists -- what `main.RiiCpp.__setstate__` of matsui528/rii casts from."""
        return (self._codewords.tolist(), bool(self.verbose), self.coarse_centers, self.flattened_codes, self.posting_lists)

    def dumps_reference(self):
        """Pickle bytes that the UNMODIFIED reference loads with `pickle.loads` into its own `main.RiiCpp`."""
        return reference_pickle_bytes(*self.to_reference_state())

    # ---- flat, memory-mappable state for large indexes -----------------------------
    def save_flat(self, path):
        """Directory of raw arrays + meta.json.  The code table is written through a memory map (no second host copy),
        so 1e8-1e9 codes need only page cache."""
        import json
        import os
        os.makedirs(path, exist_ok=True)
        np.save(os.path.join(path, "codewords.npy"), self._codewords)
        N, M, nlist = self.N, self.M, self.nlist
        codes = np.lib.format.open_memmap(os.path.join(path, "codes.npy"), mode="w+", dtype=np.uint8, shape=(N, M))
        if N:
            check(_capi.lib().rii_copy_codes(self._h, _ptr(codes, C.c_uint8)))
        codes.flush()
        del codes
        np.save(os.path.join(path, "coarse_centers.npy"), self.coarse_centers_array())
        offsets, ids = self.posting_lists_csr()
        np.save(os.path.join(path, "posting_offsets.npy"), offsets)
        np.save(os.path.join(path, "posting_ids.npy"), ids)
        with open(os.path.join(path, "meta.json"), "w") as f:
            json.dump({"format": "rii_b200-flat-1", "N": N, "M": M, "Ks": self.Ks, "Ds": self.Ds, "nlist": nlist,
                       "verbose": bool(self.verbose), "l2_variant": self._l2_variant}, f)

    @classmethod
    def load_flat(cls, path, device=0):
        """Load a save_flat() directory: the arrays are memory-mapped, never materialised on the host heap."""
        import json
        import os
        meta = json.load(open(os.path.join(path, "meta.json")))
        assert meta["format"] == "rii_b200-flat-1"
        e = cls(np.load(os.path.join(path, "codewords.npy")), meta["verbose"], device=device, l2_variant=meta["l2_variant"])
        codes = np.load(os.path.join(path, "codes.npy"), mmap_mode="r")  # pages are read as the copy engine pulls them
        centers = np.ascontiguousarray(np.load(os.path.join(path, "coarse_centers.npy")), np.uint8).reshape(-1, meta["M"])
        offsets = np.ascontiguousarray(np.load(os.path.join(path, "posting_offsets.npy")), np.int64)
        ids = np.load(os.path.join(path, "posting_ids.npy"), mmap_mode="r")
        check(_capi.lib().rii_set_state(e._h, _ptr(centers, C.c_uint8), centers.shape[0], _ptr(codes, C.c_uint8), meta["N"],
                                        _ptr(offsets, C.c_int64), _ptr(ids, C.c_int32)))
        return e

    def encode(self, vecs):
        """GPU PQ encoder: float32 (n, D) -> uint8 (n, M) (nearest codeword per sub-space, first minimum wins)."""
        X = np.ascontiguousarray(vecs, np.float32)
        if X.ndim != 2 or X.shape[1] != self.M * self.Ds:
            raise ValueError("vecs must have shape (n, M * Ds)")
        out = np.empty((X.shape[0], self.M), np.uint8)
        check(_capi.lib().rii_encode(self._h, _ptr(X, C.c_float), X.shape[0], _ptr(out, C.c_uint8)))
        return out

    def set_rotation(self, R):
        """OPQ rotation applied on the device to every query (q @ R); None switches it off."""
        if R is None:
            check(_capi.lib().rii_set_rotation(self._h, None))
            return
        R = np.ascontiguousarray(R, np.float32)
        D = self.M * self.Ds
        if R.shape != (D, D):
            raise ValueError("R must have shape (D, D)")
        check(_capi.lib().rii_set_rotation(self._h, _ptr(R, C.c_float)))

    def set_option(self, name, value):
        check(_capi.lib().rii_set_option(self._h, name.encode(), int(value)))

    # ---- building blocks (parity tests) -------------------------------------------------------
    def dtable(self, queries):
        Q = np.ascontiguousarray(queries, np.float32).reshape(-1, self.M * self.Ds)
        out = np.empty((Q.shape[0], self.M, self.Ks), np.float32)
        check(_capi.lib().rii_dtable(self._h, _ptr(Q, C.c_float), Q.shape[0], _ptr(out, C.c_float)))
        return out

    def adist_all(self, query):
        q = np.ascontiguousarray(query, np.float32)
        out = np.empty(self.N, np.float32)
        check(_capi.lib().rii_adist_all(self._h, _ptr(q, C.c_float), _ptr(out, C.c_float)))
        return out

    def assign(self, codes, centers, return_dist=False):
        codes = np.ascontiguousarray(codes, np.uint8)
        centers = np.ascontiguousarray(centers, np.uint8)
        a = np.empty(codes.shape[0], np.int32)
        d = np.empty(codes.shape[0], np.float32)
        check(_capi.lib().rii_assign(self._h, _ptr(codes, C.c_uint8), codes.shape[0], _ptr(centers, C.c_uint8),
                                     centers.shape[0], _ptr(a, C.c_int32), _ptr(d, C.c_float)))
        return (a, d) if return_dist else a

    def sym_matrices(self):
        out = np.empty((self.M, self.Ks, self.Ks), np.float32)
        check(_capi.lib().rii_sym_matrices(self._h, _ptr(out, C.c_float)))
        return out


def reference_pickle_bytes(codewords, verbose, coarse_centers, flattened_codes, posting_lists):
    """A protocol-2 pickle stream of the reference's class `main.RiiCpp` (pybind11 pickles through copyreg.__newobj__ +
    __setstate__, src/main.cpp:35-54) carrying the given 5-tuple.  Written opcode by opcode because the class `main.RiiCpp`
    need not be importable where the index lives."""
    import pickle
    state = pickle.dumps((codewords, bool(verbose), coarse_centers, flattened_codes, posting_lists), protocol=2)
    assert state[:2] == b"\x80\x02" and state[-1:] == b"."
    #       PROTO 2      GLOBAL main.RiiCpp   EMPTY_TUPLE NEWOBJ   <state>      BUILD STOP
    return b"\x80\x02" + b"cmain\nRiiCpp\n" + b")" + b"\x81" + state[2:-1] + b"b" + b"."
