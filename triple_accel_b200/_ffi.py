"""ctypes binding of libtriple_accel_b200.so (the C ABI declared in include/triple_accel_b200.h).

The library is the product; this module only loads it and declares signatures.  There is no fallback: if the
shared object is missing, or no CUDA device is usable, calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtriple_accel_b200.so")

TA_NONE = 0xFFFFFFFF
TA_OK = 0
TA_ERR_CUDA = -1
TA_ERR_LEN_MISMATCH = -2
TA_ERR_BAD_COSTS = -3
TA_ERR_BAD_ARG = -4
TA_ERR_TOO_LARGE = -5
TA_ERR_NOMEM = -6
TA_ERR_NUL_BYTE = -7
TA_SEARCH_ALL = 0
TA_SEARCH_BEST = 1


class ta_costs(C.Structure):
    _fields_ = [("mismatch", C.c_uint8), ("gap", C.c_uint8), ("start_gap", C.c_uint8), ("transpose", C.c_uint8)]


class ta_edit(C.Structure):
    _fields_ = [("edit", C.c_uint32), ("count", C.c_uint32)]


class ta_match(C.Structure):
    _fields_ = [("start", C.c_uint64), ("end", C.c_uint64), ("k", C.c_uint32), ("_pad", C.c_uint32)]


# every symbol include/triple_accel_b200.h declares: name -> (restype, argtypes)
_vp, _sz, _u32, _int = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int
SIGNATURES = {
    "ta_abi_version": (_int, []),
    "ta_init": (_int, [_int, C.POINTER(_vp)]),
    "ta_init_multi": (_int, [C.POINTER(_int), _int, C.POINTER(_vp)]),
    "ta_device_count": (_int, [_vp]),
    "ta_shard_bounds": (_int, [_vp, _vp, _sz, _int, _vp]),
    "ta_multi_uses_nccl": (_int, [_vp]),
    "ta_multi_needle_broadcasts": (C.c_uint64, [_vp]),
    "ta_trim": (None, []),
    "ta_set_length_hint": (_int, [_vp, _int]),
    "ta_shutdown": (None, [_vp]),
    "ta_strerror": (C.c_char_p, [_int]),
    "ta_last_error": (C.c_char_p, [_vp]),
    "ta_device": (_int, [_vp]),
    "ta_launch_count": (C.c_uint64, [_vp]),
    "ta_host_alloc": (_vp, [_sz]),
    "ta_host_free": (None, [_vp]),
    "ta_free": (None, [_vp]),
    "ta_costs_valid": (_int, [ta_costs]),
    "ta_costs_valid_search": (_int, [ta_costs]),
    "ta_hamming_batch": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ta_levenshtein_k_batch": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _u32, ta_costs, _vp]),
    "ta_levenshtein_exp_batch": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, ta_costs, _vp]),
    "ta_levenshtein_k_trace_batch": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _u32, ta_costs, _vp,
                                            C.POINTER(C.POINTER(ta_edit)), C.POINTER(C.POINTER(C.c_uint64))]),
    "ta_levenshtein_exp_trace_batch": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, ta_costs, _vp,
                                              C.POINTER(C.POINTER(ta_edit)), C.POINTER(C.POINTER(C.c_uint64))]),
    "ta_levenshtein_search_batch": (_int, [_vp, _vp, _sz, _vp, _vp, _sz, _u32, _int, ta_costs, _int,
                                           C.POINTER(C.POINTER(ta_match)), C.POINTER(C.POINTER(C.c_uint64))]),
    "ta_search_default_k": (_u32, [_sz]),
    "ta_hamming_search_batch": (_int, [_vp, _vp, _sz, _vp, _vp, _sz, _u32, _int,
                                       C.POINTER(C.POINTER(ta_match)), C.POINTER(C.POINTER(C.c_uint64))]),
    "ta_hamming_search_naive_batch": (_int, [_vp, _vp, _sz, _vp, _vp, _sz, _u32, _int,
                                             C.POINTER(C.POINTER(ta_match)), C.POINTER(C.POINTER(C.c_uint64))]),
    "ta_hamming_batch_dev": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "ta_hamming_batch_dev_len": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, C.c_uint32, _vp, _vp]),
    "ta_levenshtein_k_batch_dev": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _u32, ta_costs, _u32, _vp, _vp]),
    "ta_levenshtein_exp_batch_dev": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, ta_costs, _u32, _vp, _vp]),
    "ta_levenshtein_search_batch_dev": (_int, [_vp, _vp, _sz, _vp, _vp, _sz, C.c_uint64, _u32, _int, ta_costs, _int,
                                               C.POINTER(C.POINTER(ta_match)), C.POINTER(C.POINTER(C.c_uint64)),
                                               _vp]),
    "ta_dev_status": (_int, [_vp, _vp]),
    "ta_hamming": (_int, [_vp, C.c_char_p, _sz, C.c_char_p, _sz, C.POINTER(_u32)]),
    "ta_levenshtein_simd_k_with_opts": (_int, [_vp, C.c_char_p, _sz, C.c_char_p, _sz, _u32, ta_costs,
                                                C.POINTER(_u32)]),
    "ta_levenshtein_exp_with_opts": (_int, [_vp, C.c_char_p, _sz, C.c_char_p, _sz, ta_costs, C.POINTER(_u32)]),
}

_lib = None


def load():
    """Load the shared library (raises OSError with a build hint if it is missing)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C triple_accel_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
