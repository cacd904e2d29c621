"""Deterministic synthetic inputs for the BASELINE configs, numpy only.

Modelled on the reference's bench generators (benches/rand_benchmarks.rs:126-260: random strings, then a
bounded number of random edits), seeded with 1234 like the reference benches (benches/rand_benchmarks.rs:8).
Everything is returned in the CSR layout of the C ABI: (bytes uint8[], offsets uint64[n+1]).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_TOOLS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools")
_native = None


def native():
    """tools/libta_synth.so (tools/ta_synth.c): every unit draws from its own splitmix64 stream keyed by (seed, global
    unit index) -- 1 M pairs are 1 M different edit scripts, and rank r of N can make units [lo, hi) of ONE batch."""
    global _native
    if _native is None:
        so, src = os.path.join(_TOOLS, "libta_synth.so"), os.path.join(_TOOLS, "ta_synth.c")
        if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
            subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-pthread", "-o", so, src])
        L = C.CDLL(so)
        u64, u32, vp, i = C.c_uint64, C.c_uint32, C.c_void_p, C.c_int
        L.synth_pair_lengths.restype = None
        L.synth_pair_lengths.argtypes = [u64, u64, u64, u32, u32, u32, i, i, u32, vp, vp, i]
        L.synth_pair_fill.restype = None
        L.synth_pair_fill.argtypes = [u64, u64, u64, u32, u32, u32, i, i, u32, vp, vp, vp, vp, i]
        L.synth_haystacks.restype = None
        L.synth_haystacks.argtypes = [u64, u64, u64, u32, vp, u32, u32, u32, vp, i]
        _native = L
    return _native


def _threads():
    return max(1, min(32, len(os.sched_getaffinity(0))))


def _vp(x):
    return x.ctypes.data_as(C.c_void_p)


def edited_pairs(n, len_lo, len_hi, max_edits, seed=1234, first=0, exact_edits=False, allow_swap=False, alphabet=256,
                 threads=None):
    """Set M, one independent edit script per pair (tools/ta_synth.c): |a| ~ U[len_lo, len_hi], b = a after
    e ~ U[0, max_edits] (or exactly max_edits) random substitutions / insertions / deletions (/ adjacent swaps).
    Units first .. first + n of the batch defined by `seed` (same bytes whatever n, first or the thread count)."""
    L = native()
    threads = threads or _threads()
    la, lb = np.empty(n, np.uint32), np.empty(n, np.uint32)
    args = (seed, first, n, len_lo, len_hi, max_edits, int(exact_edits), int(allow_swap), alphabet)
    L.synth_pair_lengths(*args, _vp(la), _vp(lb), threads)
    a_off, b_off = np.zeros(n + 1, np.uint64), np.zeros(n + 1, np.uint64)
    np.cumsum(la, dtype=np.uint64, out=a_off[1:])
    np.cumsum(lb, dtype=np.uint64, out=b_off[1:])
    a, b = np.empty(int(a_off[-1]), np.uint8), np.empty(int(b_off[-1]), np.uint8)
    L.synth_pair_fill(*args, _vp(a), _vp(a_off), _vp(b), _vp(b_off), threads)
    return a, a_off, b, b_off


def planted_haystacks(n, hay_len, needle, plant_frac=0.01, max_edits=3, seed=1234, first=0, threads=None):
    """cfg 4 haystacks first .. first + n (bytes 1..255; the needle, mutated by <= max_edits edits, planted in
    plant_frac of them), independent per haystack like edited_pairs."""
    L = native()
    needle = np.ascontiguousarray(needle, np.uint8)
    hay = np.empty(n * hay_len, np.uint8)
    L.synth_haystacks(seed, first, n, hay_len, _vp(needle), len(needle), int(round(plant_frac * 1e6)), max_edits,
                      _vp(hay), threads or _threads())
    return hay, fixed_offsets(n, hay_len)


def _rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def fixed_offsets(n, length):
    return (np.arange(n + 1, dtype=np.uint64) * np.uint64(length))


def hamming_pairs(n, length, max_mut=6, seed=1234):
    """cfg 1: b = a with 0..max_mut positions overwritten by a different byte."""
    g = _rng(seed)
    a = g.integers(0, 256, size=(n, length), dtype=np.uint8)
    b = a.copy()
    nm = g.integers(0, max_mut + 1, size=n)
    for j in range(max_mut):
        rows = np.nonzero(nm > j)[0]
        pos = g.integers(0, length, size=len(rows))
        b[rows, pos] = a[rows, pos] + g.integers(1, 256, size=len(rows), dtype=np.uint8)
    off = fixed_offsets(n, length)
    return a.reshape(-1), off, b.reshape(-1), off.copy()


def _apply_edits(a_row, n_edits, g, allow_swap):
    s = bytearray(a_row.tobytes())
    for _ in range(n_edits):
        kind = g.integers(0, 4 if allow_swap else 3)
        if kind == 0 and s:  # substitute
            p = int(g.integers(0, len(s)))
            s[p] = (s[p] + int(g.integers(1, 256))) & 0xFF
        elif kind == 1:  # insert
            p = int(g.integers(0, len(s) + 1))
            s.insert(p, int(g.integers(0, 256)))
        elif kind == 2 and s:  # delete
            p = int(g.integers(0, len(s)))
            del s[p]
        elif kind == 3 and len(s) > 1:  # swap adjacent
            p = int(g.integers(0, len(s) - 1))
            s[p], s[p + 1] = s[p + 1], s[p]
    return bytes(s)


def mutated_pairs(n, length, max_edits, seed=1234, exact_edits=False, allow_swap=False, templates=4096):
    """cfg 2/3/5 'matching' set M: b = a after e random edits (e ~ U[0, max_edits], or exactly max_edits).

    Editing a million strings in pure Python is slow, so `templates` distinct (a, b) pairs are generated and
    tiled with a per-copy XOR mask on both strings (XOR with a constant byte preserves every edit distance),
    which keeps the batch free of byte-identical repeats.
    """
    g = _rng(seed)
    t = min(templates, n)
    a_t = g.integers(0, 256, size=(t, length), dtype=np.uint8)
    b_t = []
    for r in range(t):
        e = max_edits if exact_edits else int(g.integers(0, max_edits + 1))
        b_t.append(_apply_edits(a_t[r], e, g, allow_swap))
    reps = (n + t - 1) // t
    masks = g.integers(0, 256, size=reps, dtype=np.uint8)
    a_parts, b_parts, b_lens = [], [], []
    b_t_arr = [np.frombuffer(x, np.uint8) for x in b_t]
    b_cat = np.concatenate(b_t_arr) if t else np.zeros(0, np.uint8)
    lens_t = np.array([len(x) for x in b_t], np.uint64)
    for r in range(reps):
        cnt = min(t, n - r * t)
        a_parts.append(a_t[:cnt] ^ masks[r])
        nb = int(lens_t[:cnt].sum())
        b_parts.append(b_cat[:nb] ^ masks[r])
        b_lens.append(lens_t[:cnt])
    a = np.concatenate(a_parts).reshape(-1)
    b = np.concatenate(b_parts)
    b_off = np.zeros(n + 1, np.uint64)
    b_off[1:] = np.cumsum(np.concatenate(b_lens))
    return a, fixed_offsets(n, length), b, b_off


def ragged_mutated_pairs(n, min_len, max_len, max_edits, seed=1234, templates=4096):
    """set M with ragged lengths: |a| ~ U[min_len, max_len], b = a after e ~ U[0, max_edits] random edits.
    Same template + XOR-mask construction as mutated_pairs; consecutive pairs have unrelated lengths."""
    g = _rng(seed)
    t = min(templates, n)
    a_t, b_t = [], []
    for r in range(t):
        la = int(g.integers(min_len, max_len + 1))
        row = g.integers(0, 256, size=la, dtype=np.uint8)
        a_t.append(row)
        b_t.append(np.frombuffer(_apply_edits(row, int(g.integers(0, max_edits + 1)), g, False), np.uint8))
    reps = (n + t - 1) // t
    masks = g.integers(0, 256, size=reps, dtype=np.uint8)
    a_cat, b_cat = np.concatenate(a_t), np.concatenate(b_t)
    a_len = np.array([len(x) for x in a_t], np.uint64)
    b_len = np.array([len(x) for x in b_t], np.uint64)
    a_parts, b_parts, al, bl = [], [], [], []
    for r in range(reps):
        cnt = min(t, n - r * t)
        a_parts.append(a_cat[:int(a_len[:cnt].sum())] ^ masks[r])
        b_parts.append(b_cat[:int(b_len[:cnt].sum())] ^ masks[r])
        al.append(a_len[:cnt])
        bl.append(b_len[:cnt])
    a_off = np.zeros(n + 1, np.uint64)
    b_off = np.zeros(n + 1, np.uint64)
    a_off[1:] = np.cumsum(np.concatenate(al))
    b_off[1:] = np.cumsum(np.concatenate(bl))
    return np.concatenate(a_parts), a_off, np.concatenate(b_parts), b_off


def random_pairs(n, length, seed=1234):
    """set R: independent uniform byte strings of equal length (almost every pair is farther apart than k)."""
    g = _rng(seed)
    a = g.integers(0, 256, size=n * length, dtype=np.uint8)
    b = g.integers(0, 256, size=n * length, dtype=np.uint8)
    off = fixed_offsets(n, length)
    return a, off, b, off.copy()


def needle_haystacks(n, hay_len, needle_len=32, plant_frac=0.01, max_edits=3, seed=1234):
    """cfg 4: needle of bytes 1..255; haystacks of bytes 1..255; the needle (mutated by <= max_edits edits) is
    planted in plant_frac of the haystacks at a random offset (model: benches/rand_benchmarks.rs:175-198)."""
    g = _rng(seed)
    needle = g.integers(1, 256, size=needle_len, dtype=np.uint8)
    hay = g.integers(1, 256, size=(n, hay_len), dtype=np.uint8)
    n_plant = int(round(n * plant_frac))
    rows = g.choice(n, size=n_plant, replace=False) if n_plant else np.zeros(0, np.int64)
    for r in rows:
        mut = np.frombuffer(_apply_edits(needle, int(g.integers(0, max_edits + 1)), g, False), np.uint8)
        mut = np.where(mut == 0, 1, mut).astype(np.uint8)
        p = int(g.integers(0, hay_len - len(mut) + 1))
        hay[r, p:p + len(mut)] = mut
    return needle, hay.reshape(-1), fixed_offsets(n, hay_len)
