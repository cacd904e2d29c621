"""ctypes binding of oracle/libta_oracle.so -- the CPU restatement of the reference's scalar functions.

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
NONE = 0xFFFFFFFF


class Costs(C.Structure):
    _fields_ = [("mismatch", C.c_uint8), ("gap", C.c_uint8), ("start_gap", C.c_uint8), ("transpose", C.c_uint8)]


class Match(C.Structure):
    _fields_ = [("start", C.c_uint64), ("end", C.c_uint64), ("k", C.c_uint32), ("_pad", C.c_uint32)]


class Edit(C.Structure):
    _fields_ = [("edit", C.c_uint32), ("count", C.c_uint32)]


LEVENSHTEIN_COSTS = (1, 1, 0, 0)
RDAMERAU_COSTS = (1, 1, 0, 1)

_lib = None


def build():
    so = os.path.join(ORACLE_DIR, "libta_oracle.so")
    src = [os.path.join(ORACLE_DIR, f) for f in ("ta_oracle.c", "ta_ref_avx2.c", "ta_oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return so


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        u8p, u64p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
        L.orc_hamming_naive.restype = C.c_int64
        L.orc_hamming_naive.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.orc_levenshtein_naive_with_opts.restype = C.c_uint32
        L.orc_levenshtein_naive_with_opts.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, Costs,
                                                      C.POINTER(C.POINTER(Edit)), C.POINTER(C.c_size_t)]
        L.orc_levenshtein_naive_k_with_opts.restype = C.c_uint32
        L.orc_levenshtein_naive_k_with_opts.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_uint32,
                                                        Costs, C.POINTER(C.POINTER(Edit)), C.POINTER(C.c_size_t)]
        L.orc_levenshtein_exp_with_opts.restype = C.c_uint32
        L.orc_levenshtein_exp_with_opts.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, Costs]
        L.orc_levenshtein_search_naive_with_opts.restype = C.c_int64
        L.orc_levenshtein_search_naive_with_opts.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
                                                             C.c_uint32, C.c_int, Costs, C.c_int,
                                                             C.POINTER(C.POINTER(Match))]
        for nm in ("orc_hamming_search_naive_with_opts", "orc_hamming_search_with_opts"):
            f = getattr(L, nm)
            f.restype = C.c_int64
            f.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_uint32, C.c_int,
                          C.POINTER(C.POINTER(Match))]
        L.orc_hamming_search_batch.restype = C.c_int64
        L.orc_hamming_search_batch.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32,
                                               C.c_int, C.POINTER(C.POINTER(Match)), C.c_void_p, C.c_int]
        L.orc_search_default_k.restype = C.c_uint32
        L.orc_search_default_k.argtypes = [C.c_size_t]
        L.orc_costs_valid.argtypes = [Costs]
        L.orc_costs_valid_search.argtypes = [Costs]
        L.orc_hamming_batch.restype = None
        L.orc_hamming_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        L.orc_levenshtein_k_batch.restype = None
        L.orc_levenshtein_k_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32,
                                              Costs, C.c_void_p, C.c_int]
        L.orc_levenshtein_exp_batch.restype = None
        L.orc_levenshtein_exp_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, Costs,
                                                C.c_void_p, C.c_int]
        L.orc_levenshtein_search_batch.restype = C.c_int64
        L.orc_levenshtein_search_batch.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                                   C.c_uint32, C.c_int, Costs, C.c_int, C.POINTER(C.POINTER(Match)),
                                                   C.c_void_p, C.c_int]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_max_threads.restype = C.c_int
        # AVX2 restatement of the reference's SIMD path (CPU baseline only, oracle/ta_ref_avx2.c)
        L.orc_simd_available.restype = C.c_int
        L.orc_simd_covers.restype = C.c_int
        L.orc_simd_covers.argtypes = [C.c_size_t, C.c_size_t, C.c_uint32, Costs]
        L.orc_levenshtein_simd_k_with_opts.restype = C.c_uint32
        L.orc_levenshtein_simd_k_with_opts.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_uint32,
                                                       Costs, C.POINTER(C.c_int)]
        L.orc_levenshtein_simd_exp_with_opts.restype = C.c_uint32
        L.orc_levenshtein_simd_exp_with_opts.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, Costs]
        L.orc_hamming_simd.restype = C.c_int64
        L.orc_hamming_simd.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.orc_levenshtein_simd_k_batch.restype = None
        L.orc_levenshtein_simd_k_batch.argtypes = L.orc_levenshtein_k_batch.argtypes
        L.orc_levenshtein_simd_exp_batch.restype = None
        L.orc_levenshtein_simd_exp_batch.argtypes = L.orc_levenshtein_exp_batch.argtypes
        L.orc_hamming_simd_batch.restype = None
        L.orc_hamming_simd_batch.argtypes = L.orc_hamming_batch.argtypes
        L.orc_levenshtein_search_simd_with_opts.restype = C.c_int64
        L.orc_levenshtein_search_simd_with_opts.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_uint32,
                                                            C.c_int, Costs, C.c_int, C.POINTER(C.POINTER(Match)),
                                                            C.POINTER(C.c_int)]
        L.orc_levenshtein_search_simd_batch.restype = C.c_int64
        L.orc_levenshtein_search_simd_batch.argtypes = L.orc_levenshtein_search_batch.argtypes
        _lib = L
    return _lib


def _edits(ep, n):
    out = [(ep[i].edit, ep[i].count) for i in range(n.value)]
    if ep:
        lib().orc_free(ep)
    return out


def hamming_naive(a: bytes, b: bytes):
    r = lib().orc_hamming_naive(a, len(a), b, len(b))
    if r < 0:
        raise AssertionError("hamming: length mismatch (reference panics, src/hamming.rs:38)")
    return r


def levenshtein_naive_with_opts(a, b, trace_on=False, costs=LEVENSHTEIN_COSTS):
    if trace_on:
        ep, n = C.POINTER(Edit)(), C.c_size_t()
        d = lib().orc_levenshtein_naive_with_opts(a, len(a), b, len(b), Costs(*costs), C.byref(ep), C.byref(n))
        return d, _edits(ep, n)
    return lib().orc_levenshtein_naive_with_opts(a, len(a), b, len(b), Costs(*costs), None, None), None


def levenshtein_naive_k_with_opts(a, b, k, trace_on=False, costs=LEVENSHTEIN_COSTS):
    """Returns None, or (dist, edits|None)."""
    if trace_on:
        ep, n = C.POINTER(Edit)(), C.c_size_t()
        d = lib().orc_levenshtein_naive_k_with_opts(a, len(a), b, len(b), k, Costs(*costs), C.byref(ep), C.byref(n))
        if d == NONE:
            return None
        return d, _edits(ep, n)
    d = lib().orc_levenshtein_naive_k_with_opts(a, len(a), b, len(b), k, Costs(*costs), None, None)
    return None if d == NONE else (d, None)


def levenshtein_exp_with_opts(a, b, costs=LEVENSHTEIN_COSTS):
    return lib().orc_levenshtein_exp_with_opts(a, len(a), b, len(b), Costs(*costs))


def search_default_k(n):
    return lib().orc_search_default_k(n)


def levenshtein_search_naive_with_opts(needle, haystack, k, search_type=0, costs=LEVENSHTEIN_COSTS, anchored=False):
    mp = C.POINTER(Match)()
    n = lib().orc_levenshtein_search_naive_with_opts(needle, len(needle), haystack, len(haystack), k, search_type,
                                                     Costs(*costs), int(anchored), C.byref(mp))
    if n < 0:
        raise AssertionError("check_search failed (reference panics, src/levenshtein.rs:69)")
    out = [(mp[i].start, mp[i].end, mp[i].k) for i in range(n)]
    if mp:
        lib().orc_free(mp)
    return out


def hamming_search_naive_with_opts(needle, haystack, k, search_type=0, public_entry=False):
    """scalar routine (src/hamming.rs:96-146); public_entry=True adds the checks of hamming_search_simd_with_opts"""
    mp = C.POINTER(Match)()
    fn = lib().orc_hamming_search_with_opts if public_entry else lib().orc_hamming_search_naive_with_opts
    n = fn(needle, len(needle), haystack, len(haystack), k, search_type, C.byref(mp))
    if n == -2:
        raise AssertionError("No zero/null bytes allowed in the string! (src/lib.rs:240)")
    out = [(mp[i].start, mp[i].end, mp[i].k) for i in range(n)]
    if mp:
        lib().orc_free(mp)
    return out


def hamming_search_batch(needle, hay, hay_off, k, search_type=0, threads=1):
    n = len(hay_off) - 1
    mp = C.POINTER(Match)()
    moff = np.zeros(n + 1, np.uint64)
    total = lib().orc_hamming_search_batch(bytes(needle), len(needle), _p(hay), _p(hay_off), n, k, search_type,
                                           C.byref(mp), _p(moff), threads)
    if total == -2:
        raise AssertionError("No zero/null bytes allowed in the string! (src/lib.rs:240)")
    arr = np.zeros((total, 3), np.uint64)
    if total:
        raw = np.ctypeslib.as_array(C.cast(mp, C.POINTER(C.c_uint64)), shape=(total, 3)).copy()
        arr[:, 0], arr[:, 1] = raw[:, 0], raw[:, 1]
        arr[:, 2] = raw[:, 2] & 0xFFFFFFFF
    if mp:
        lib().orc_free(mp)
    return arr, moff


# ---- batch (CSR numpy arrays) --------------------------------------------------------------------
def _p(x):
    return x.ctypes.data_as(C.c_void_p)


def hamming_batch(a, a_off, b, b_off, threads=1):
    n = len(a_off) - 1
    out = np.empty(n, np.uint32)
    lib().orc_hamming_batch(_p(a), _p(a_off), _p(b), _p(b_off), n, _p(out), threads)
    return out


def levenshtein_k_batch(a, a_off, b, b_off, k, costs=LEVENSHTEIN_COSTS, threads=1):
    n = len(a_off) - 1
    out = np.empty(n, np.uint32)
    lib().orc_levenshtein_k_batch(_p(a), _p(a_off), _p(b), _p(b_off), n, k, Costs(*costs), _p(out), threads)
    return out


def levenshtein_exp_batch(a, a_off, b, b_off, costs=LEVENSHTEIN_COSTS, threads=1):
    n = len(a_off) - 1
    out = np.empty(n, np.uint32)
    lib().orc_levenshtein_exp_batch(_p(a), _p(a_off), _p(b), _p(b_off), n, Costs(*costs), _p(out), threads)
    return out


def levenshtein_search_batch(needle, hay, hay_off, k, search_type=0, costs=LEVENSHTEIN_COSTS, anchored=False,
                             threads=1):
    """Returns (matches[total,3] uint64 array of start,end,k ; match_off[n+1])."""
    n = len(hay_off) - 1
    mp = C.POINTER(Match)()
    moff = np.zeros(n + 1, np.uint64)
    total = lib().orc_levenshtein_search_batch(bytes(needle), len(needle), _p(hay), _p(hay_off), n, k, search_type,
                                               Costs(*costs), int(anchored), C.byref(mp), _p(moff), threads)
    if total < 0:
        raise AssertionError("check_search failed")
    arr = np.zeros((total, 3), np.uint64)
    if total:
        raw = np.ctypeslib.as_array(C.cast(mp, C.POINTER(C.c_uint64)), shape=(total, 3)).copy()
        arr[:, 0], arr[:, 1] = raw[:, 0], raw[:, 1]
        arr[:, 2] = raw[:, 2] & 0xFFFFFFFF
    if mp:
        lib().orc_free(mp)
    return arr, moff


def max_threads():
    return lib().orc_max_threads()


# ---- AVX2 restatement of the reference's SIMD path (CPU baseline; NOT the parity oracle) -----------------------
def simd_available():
    return bool(lib().orc_simd_available())


def levenshtein_simd_k_with_opts(a, b, k, costs=LEVENSHTEIN_COSTS):
    """(distance | None, covered): covered is False when the reference would use a wider Jewel type than Avx1x32x8"""
    cov = C.c_int(0)
    d = lib().orc_levenshtein_simd_k_with_opts(a, len(a), b, len(b), k, Costs(*costs), C.byref(cov))
    return (None if d == NONE else d), bool(cov.value)


def hamming_simd(a, b):
    return lib().orc_hamming_simd(a, len(a), b, len(b))


def levenshtein_search_simd_with_opts(needle, haystack, k, search_type=0, costs=LEVENSHTEIN_COSTS, anchored=False):
    """([(start, end, k), ...], covered) from the restatement of the reference's AVX2 search core"""
    mp, cov = C.POINTER(Match)(), C.c_int(0)
    n = lib().orc_levenshtein_search_simd_with_opts(needle, len(needle), haystack, len(haystack), k, search_type,
                                                    Costs(*costs), int(anchored), C.byref(mp), C.byref(cov))
    if n < 0:
        raise AssertionError("check_search failed (reference panics, src/levenshtein.rs:69)")
    out = [(mp[i].start, mp[i].end, mp[i].k) for i in range(n)]
    if mp:
        lib().orc_free(mp)
    return out, bool(cov.value)


def levenshtein_search_simd_batch(needle, hay, hay_off, k, search_type=0, costs=LEVENSHTEIN_COSTS, anchored=False,
                                  threads=1):
    n = len(hay_off) - 1
    mp = C.POINTER(Match)()
    moff = np.zeros(n + 1, np.uint64)
    total = lib().orc_levenshtein_search_simd_batch(bytes(needle), len(needle), _p(hay), _p(hay_off), n, k, search_type,
                                                    Costs(*costs), int(anchored), C.byref(mp), _p(moff), threads)
    if total < 0:
        raise AssertionError("check_search failed")
    arr = np.zeros((total, 3), np.uint64)
    if total:
        raw = np.ctypeslib.as_array(C.cast(mp, C.POINTER(C.c_uint64)), shape=(total, 3)).copy()
        arr[:, 0], arr[:, 1] = raw[:, 0], raw[:, 1]
        arr[:, 2] = raw[:, 2] & 0xFFFFFFFF
    if mp:
        lib().orc_free(mp)
    return arr, moff


def levenshtein_simd_k_batch(a, a_off, b, b_off, k, costs=LEVENSHTEIN_COSTS, threads=1):
    n = len(a_off) - 1
    out = np.empty(n, np.uint32)
    lib().orc_levenshtein_simd_k_batch(_p(a), _p(a_off), _p(b), _p(b_off), n, k, Costs(*costs), _p(out), threads)
    return out


def levenshtein_simd_exp_batch(a, a_off, b, b_off, costs=LEVENSHTEIN_COSTS, threads=1):
    n = len(a_off) - 1
    out = np.empty(n, np.uint32)
    lib().orc_levenshtein_simd_exp_batch(_p(a), _p(a_off), _p(b), _p(b_off), n, Costs(*costs), _p(out), threads)
    return out


def hamming_simd_batch(a, a_off, b, b_off, threads=1):
    n = len(a_off) - 1
    out = np.empty(n, np.uint32)
    lib().orc_hamming_simd_batch(_p(a), _p(a_off), _p(b), _p(b_off), n, _p(out), threads)
    return out
