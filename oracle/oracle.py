"""ctypes front-end of oracle/liboracle.so (the CPU restatement, oracle/rii_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  Nothing under rii_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile liboracle.so (and, when /root/reference exists, oracle/_ref) with oracle/Makefile."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "rii_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        f32p, u8p, i64p, i32p = (C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_int32))
        L.orc_l2sqr.restype = C.c_float
        L.orc_l2sqr.argtypes = [f32p, f32p, C.c_int, C.c_int]
        L.orc_dtable.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
        L.orc_adist_all.argtypes = [f32p, u8p, C.c_int64, C.c_int, C.c_int, f32p]
        L.orc_query_linear.restype = C.c_int64
        L.orc_query_linear.argtypes = [f32p, u8p, C.c_int64, C.c_int, C.c_int, C.c_int, i64p, C.c_int64, i64p, f32p]
        L.orc_query_ivf.restype = C.c_int64
        L.orc_query_ivf.argtypes = [f32p, u8p, C.c_int64, C.c_int, C.c_int, u8p, C.c_int, i64p, i32p,
                                    C.c_int, i64p, C.c_int64, C.c_int64, i64p, f32p, i64p]
        L.orc_sym_matrices.argtypes = [f32p, C.c_int, C.c_int, C.c_int, f32p]
        L.orc_assign.argtypes = [f32p, u8p, C.c_int64, u8p, C.c_int, C.c_int, C.c_int, i32p, f32p]
        L.orc_reconfigure.argtypes = [f32p, C.c_int, C.c_int, C.c_int, u8p, C.c_int64, C.c_int, C.c_int, u8p, i32p]
        _LIB = L
    return _LIB


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def host_variant():
    """Accumulator width the reference would pick with -march=native on this host (src/distance.h:113,172,219)."""
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return 8
    if " avx512f" in flags:
        return 16
    if " avx " in flags or " avx2" in flags:
        return 8
    return 4


def l2sqr(x, y, variant=16):
    x, y = _f32(x), _f32(y)
    return float(lib().orc_l2sqr(_p(x, C.c_float), _p(y, C.c_float), x.size, variant))


def dtable(q, codewords, variant=16):
    """(M, Ks) float32 distance table of one query; SURVEY A.1 / src/rii.h:361-373."""
    cw = _f32(codewords)
    M, Ks, Ds = cw.shape
    q = _f32(q)
    assert q.shape == (M * Ds,)
    T = np.empty((M, Ks), np.float32)
    lib().orc_dtable(_p(q, C.c_float), _p(cw, C.c_float), M, Ks, Ds, variant, _p(T, C.c_float))
    return T


def adist_all(T, codes):
    T, codes = _f32(T), _u8(codes)
    N, M = codes.shape
    out = np.empty(N, np.float32)
    lib().orc_adist_all(_p(T, C.c_float), _p(codes, C.c_uint8), N, M, T.shape[1], _p(out, C.c_float))
    return out


def query_linear(T, codes, topk, tids=None):
    T, codes = _f32(T), _u8(codes)
    N, M = codes.shape
    tids = np.ascontiguousarray(tids if tids is not None else [], dtype=np.int64)
    ids = np.empty(topk, np.int64)
    dists = np.empty(topk, np.float32)
    n = lib().orc_query_linear(_p(T, C.c_float), _p(codes, C.c_uint8), N, M, T.shape[1], topk,
                               _p(tids, C.c_int64), tids.size, _p(ids, C.c_int64), _p(dists, C.c_float))
    return ids[:n], dists[:n]


def lists_to_csr(posting_lists):
    offsets = np.zeros(len(posting_lists) + 1, np.int64)
    offsets[1:] = np.cumsum([len(p) for p in posting_lists])
    ids = (np.concatenate([np.asarray(p, np.int32) for p in posting_lists]) if offsets[-1] else
           np.zeros(0, np.int32)).astype(np.int32)
    return offsets, ids


def assign_to_lists(assign, nlist):
    """Posting lists as the reference builds them (src/rii.h:356-358): ascending ids per list."""
    assign = np.asarray(assign)
    order = np.argsort(assign, kind="stable")
    counts = np.bincount(assign, minlength=nlist)
    offsets = np.zeros(nlist + 1, np.int64)
    offsets[1:] = np.cumsum(counts)
    return offsets, order.astype(np.int32)


def query_ivf(T, codes, centers, offsets, ids, topk, L, tids=None, return_ncand=False):
    T, codes, centers = _f32(T), _u8(codes), _u8(centers)
    N, M = codes.shape
    offsets = np.ascontiguousarray(offsets, np.int64)
    ids = np.ascontiguousarray(ids, np.int32)
    tids = np.ascontiguousarray(tids if tids is not None else [], dtype=np.int64)
    out_ids = np.empty(topk, np.int64)
    out_d = np.empty(topk, np.float32)
    ncand = C.c_int64(0)
    n = lib().orc_query_ivf(_p(T, C.c_float), _p(codes, C.c_uint8), N, M, T.shape[1], _p(centers, C.c_uint8),
                            centers.shape[0], _p(offsets, C.c_int64), _p(ids, C.c_int32), topk,
                            _p(tids, C.c_int64), tids.size, L, _p(out_ids, C.c_int64), _p(out_d, C.c_float),
                            C.byref(ncand))
    if return_ncand:
        return out_ids[:n], out_d[:n], ncand.value
    return out_ids[:n], out_d[:n]


def sym_matrices(codewords):
    cw = _f32(codewords)
    M, Ks, Ds = cw.shape
    Dm = np.empty((M, Ks, Ks), np.float32)
    lib().orc_sym_matrices(_p(cw, C.c_float), M, Ks, Ds, _p(Dm, C.c_float))
    return Dm


def assign(Dm, codes, centers, return_dist=False):
    Dm, codes, centers = _f32(Dm), _u8(codes), _u8(centers)
    N, M = codes.shape
    a = np.empty(N, np.int32)
    d = np.empty(N, np.float32)
    lib().orc_assign(_p(Dm, C.c_float), _p(codes, C.c_uint8), N, _p(centers, C.c_uint8), centers.shape[0], M,
                     Dm.shape[1], _p(a, C.c_int32), _p(d, C.c_float))
    return (a, d) if return_dist else a


def reconfigure(codewords, codes, nlist, iter=5):
    """-> (coarse_centers (nlist, M) uint8, assignment (N,) int32); src/rii.h:108-156."""
    cw, codes = _f32(codewords), _u8(codes)
    M, Ks, Ds = cw.shape
    N = codes.shape[0]
    centers = np.empty((nlist, M), np.uint8)
    a = np.empty(N, np.int32)
    lib().orc_reconfigure(_p(cw, C.c_float), M, Ks, Ds, _p(codes, C.c_uint8), N, nlist, iter,
                          _p(centers, C.c_uint8), _p(a, C.c_int32))
    return centers, a
