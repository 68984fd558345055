"""ctypes binding of the C ABI in include/rii_b200.h (librii_b200.so, built in-tree by rii_b200/build.py).

There is no CPU fallback: if the CUDA library is missing or fails to load this module raises, and so does
everything that imports it."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librii_b200.so")

# every symbol include/rii_b200.h declares: (restype, argtypes)
_f32p, _u8p, _i64p, _i32p, _vp = (C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int64),
                                  C.POINTER(C.c_int32), C.c_void_p)
SYMBOLS = {
    "rii_last_error": (C.c_char_p, []),
    "rii_version": (C.c_char_p, []),
    "rii_launch_count": (C.c_int64, []),
    "rii_create": (C.c_int, [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "rii_destroy": (C.c_int, [_vp]),
    "rii_add_codes": (C.c_int, [_vp, _u8p, C.c_int64, C.c_int]),
    "rii_add_codes_dev": (C.c_int, [_vp, _vp, C.c_int64, C.c_int]),
    "rii_reconfigure": (C.c_int, [_vp, C.c_int, C.c_int]),
    "rii_clear": (C.c_int, [_vp]),
    # (void pointers: the single-query wrappers pass raw addresses -- building typed ctypes pointers costs microseconds)
    "rii_query_linear": (C.c_int64, [_vp, _vp, C.c_int, _vp, C.c_int64, _vp, _vp]),
    "rii_query_ivf": (C.c_int64, [_vp, _vp, C.c_int, _vp, C.c_int64, C.c_int64, _vp, _vp]),
    "rii_query_batch": (C.c_int, [_vp, _f32p, C.c_int, C.c_int, _i64p, C.c_int64, C.c_int64, C.c_int, _i64p, _f32p,
                                  _i32p]),
    "rii_query_batch_dev": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, C.c_int64, C.c_int64, C.c_int, _vp, _vp, _vp,
                                      _vp]),
    "rii_get_N": (C.c_int64, [_vp]),
    "rii_get_nlist": (C.c_int, [_vp]),
    "rii_get_verbose": (C.c_int, [_vp]),
    "rii_set_verbose": (C.c_int, [_vp, C.c_int]),
    "rii_get_dims": (C.c_int, [_vp, _i32p, _i32p, _i32p]),
    "rii_copy_codes": (C.c_int, [_vp, _u8p]),
    "rii_copy_coarse_centers": (C.c_int, [_vp, _u8p]),
    "rii_copy_posting_lists": (C.c_int, [_vp, _i64p, _i32p]),
    "rii_set_state": (C.c_int, [_vp, _u8p, C.c_int, _u8p, C.c_int64, _i64p, _i32p]),
    "rii_dtable": (C.c_int, [_vp, _f32p, C.c_int, _f32p]),
    "rii_adist_all": (C.c_int, [_vp, _f32p, _f32p]),
    "rii_assign": (C.c_int, [_vp, _u8p, C.c_int64, _u8p, C.c_int, _i32p, _f32p]),
    "rii_sym_matrices": (C.c_int, [_vp, _f32p]),
    "rii_encode": (C.c_int, [_vp, _f32p, C.c_int64, _u8p]),
    "rii_set_rotation": (C.c_int, [_vp, _f32p]),
    "rii_set_shard": (C.c_int, [_vp, C.c_int64, C.c_int64]),
    "rii_set_coarse_centers": (C.c_int, [_vp, _u8p, C.c_int]),
    "rii_fit_coarse": (C.c_int, [_vp, _u8p, C.c_int64, C.c_int, C.c_int, _u8p]),
    "rii_copy_list_lengths": (C.c_int, [_vp, _i32p]),
    "rii_set_global_lengths": (C.c_int, [_vp, _i32p, _i32p]),
    "rii_sample_ids": (C.c_int, [C.c_int64, C.c_int, _i64p, _i64p]),
    "rii_set_lists_dev": (C.c_int, [_vp, _u8p, C.c_int, _vp]),
    "rii_reserve": (C.c_int, [_vp, C.c_int64]),
    "rii_merge_shards_packed_dev": (C.c_int, [_vp, _vp, C.c_int64, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "rii_merge_shards_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "rii_subset_begin_dev": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp]),
    "rii_subset_set_global_dev": (C.c_int, [_vp, _vp, _vp, _vp]),
    "rii_subset_query_dev": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp, _vp]),
    "rii_coarse_width": (C.c_int, [_vp, C.c_int64]),
    "rii_coarse_rank_dev": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int64, _vp, _vp]),
    "rii_query_ranked_dev": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rii_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    "rii_profile_enable": (C.c_int, [_vp, C.c_int]),
    "rii_profile_reset": (C.c_int, [_vp]),
    "rii_debug_clocks": (C.c_int, [_vp, C.c_int64, _i64p]),
    "rii_profile_get": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_double), _i64p]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("rii_b200: %s is missing - build it with `python -m rii_b200.build` "
                               "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


class RiiError(RuntimeError):
    pass


def check(rc):
    """Negative return code -> exception carrying rii_last_error().  RII_ERR_ARG maps to AssertionError-like
    ValueError (the reference asserts), everything else to RiiError."""
    if rc is not None and rc < 0:
        msg = lib().rii_last_error().decode()
        if rc == -1:
            raise ValueError(msg)
        raise RiiError("rii_b200 error %d: %s" % (rc, msg))
    return rc
