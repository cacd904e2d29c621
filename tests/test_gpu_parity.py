"""GPU parity tests: the CUDA path through the C ABI vs the CPU oracle (bit-exact) and vs the reference's own
known-answer vectors (tests/golden/kat.json).  Run on the B200 box with `pytest -m gpu`."""
import json
import os
import random

import numpy as np
import pytest

import _oracle as orc

pytestmark = pytest.mark.gpu

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))
NONE = 0xFFFFFFFF

COST_MODELS = [(1, 1, 0, 0), (1, 1, 0, 1), (1, 1, 2, 0), (2, 1, 2, 0), (2, 3, 0, 0), (3, 1, 0, 0), (2, 2, 1, 3),
               (3, 2, 0, 2), (1, 1, 1, 1), (5, 4, 3, 0)]


@pytest.fixture(scope="module")
def eng():
    import triple_accel_b200 as ta
    e = ta.Engine(0)
    yield e
    e.close()


def _ids(recs):
    return ["%s@%s" % (r["fn"], r["src"].split(" ")[0]) for r in recs]


def _pack(strs):
    from triple_accel_b200 import pack
    return pack(strs)


# ---------------------------------------------------------------------------------------------------------------
# known-answer vectors of the reference, through the single-pair C-ABI functions
HAMMING = [r for r in KAT if r["fn"] in ("hamming", "hamming_naive", "hamming_simd_parallel", "hamming_simd_movemask")]
DIST = [r for r in KAT if r["fn"] in ("levenshtein", "levenshtein_naive", "levenshtein_naive_with_opts", "rdamerau",
                                      "levenshtein_exp", "levenshtein_exp_with_opts", "rdamerau_exp",
                                      "levenshtein_naive_k", "levenshtein_simd_k", "levenshtein_naive_k_with_opts",
                                      "levenshtein_simd_k_with_opts")]
SEARCH = [r for r in KAT if r["fn"].startswith("levenshtein_search")]


@pytest.mark.parametrize("r", HAMMING, ids=_ids(HAMMING))
def test_kat_hamming(eng, r):
    # every Hamming variant of the crate (naive, words, movemask, parallel) is a name for the one GPU path
    fn = getattr(eng, r["fn"])
    assert fn(bytes.fromhex(r["a"]), bytes.fromhex(r["b"])) == r["expect"]["dist"]
    assert eng.hamming(bytes.fromhex(r["a"]), bytes.fromhex(r["b"])) == r["expect"]["dist"]


@pytest.mark.parametrize("r", DIST, ids=_ids(DIST))
def test_kat_distance(eng, r):
    import triple_accel_b200 as ta
    a, b = bytes.fromhex(r["a"]), bytes.fromhex(r["b"])
    fn = r["fn"]
    rd = fn.startswith("rdamerau")
    c = r.get("costs", [1, 1, 0, 1] if rd else [1, 1, 0, 0])
    costs = ta.EditCosts(c[0], c[1], c[2], c[3] or None)
    if fn in ("levenshtein_exp", "levenshtein_exp_with_opts", "rdamerau_exp"):
        assert eng.levenshtein_exp_with_opts(a, b, False, costs)[0] == r["expect"]["dist"]
        if r.get("trace"):
            d, edits = eng.levenshtein_exp_with_opts(a, b, True, costs)
            assert d == r["expect"]["dist"] and [list(e) for e in edits] == r["expect"]["edits"]
        return
    k = r.get("k", 0xFFFFFFFF)  # levenshtein()/rdamerau()/naive: k = u32::MAX
    res = eng.levenshtein_simd_k_with_opts(a, b, k, False, costs)
    if r["expect"].get("none"):
        assert res is None
    else:
        assert res is not None and res[0] == r["expect"]["dist"]
    # trace_on = true KATs of the k-bounded routines (tests/basic_tests.rs:395-427, 545-577 and doc-tests) and of the
    # unbounded levenshtein_naive_with_opts (tests/basic_tests.rs:163-195): same decision function, see api.py
    if r.get("trace") and fn in ("levenshtein_naive_k_with_opts", "levenshtein_simd_k_with_opts"):
        d, edits = eng.levenshtein_simd_k_with_opts(a, b, k, True, costs)
        assert d == r["expect"]["dist"] and [list(e) for e in edits] == r["expect"]["edits"]
    if fn == "levenshtein_naive_with_opts":
        d, edits = eng.levenshtein_naive_with_opts(a, b, bool(r.get("trace")), costs)
        assert d == r["expect"]["dist"]
        if r.get("trace"):
            assert [list(e) for e in edits] == r["expect"]["edits"]


@pytest.mark.parametrize("r", SEARCH, ids=_ids(SEARCH))
def test_kat_search(eng, r):
    import triple_accel_b200 as ta
    needle, hay = bytes.fromhex(r["a"]), bytes.fromhex(r["b"])
    if "k" in r:
        c = r["costs"]
        got = eng.levenshtein_search_simd_with_opts(needle, hay, r["k"], r["search_type"],
                                                    ta.EditCosts(c[0], c[1], c[2], c[3] or None), r["anchored"])
    else:
        got = eng.levenshtein_search(needle, hay)
    got = [tuple(m) for m in got]
    if r["expect"].get("first_only"):
        got = got[:1]
    assert got == [(m["start"], m["end"], m["k"]) for m in r["expect"]["matches"]]


def test_levenshtein_simd_k_str(eng):
    """src/levenshtein.rs:641-651 (doc-test :637-639) -- chars, not bytes, for non-ASCII strings"""
    assert eng.levenshtein_simd_k_str("abc", "ab", 1) == 1
    assert eng.levenshtein_simd_k_str("abc", "xyz", 2) is None
    assert eng.levenshtein_simd_k_str("nai\u0308ve caf\u00e9", "naive cafe", 5) == 2  # 1 deleted char + 1 substituted
    assert eng.levenshtein_simd_k_str("\u4f60\u597d\u4e16\u754c", "\u4f60\u597d\u4e16", 3) == 1
    many = "".join(chr(0x4e00 + i) for i in range(300))
    assert eng.levenshtein_simd_k_str(many, "x", 400) is None  # > 256 distinct chars (translate_str returns None)
    rng = random.Random(8)
    for _ in range(50):
        alpha = [chr(c) for c in (0x61, 0x62, 0xe9, 0x4f60, 0x1f600)]
        a = "".join(rng.choice(alpha) for _ in range(rng.randrange(0, 12)))
        b = "".join(rng.choice(alpha) for _ in range(rng.randrange(0, 12)))
        code = {ch: i for i, ch in enumerate(alpha)}
        want = orc.levenshtein_naive_k_with_opts(bytes(code[c] for c in a), bytes(code[c] for c in b), 6)
        assert eng.levenshtein_simd_k_str(a, b, 6) == (None if want is None else want[0])


HSEARCH = [r for r in KAT if r["fn"].startswith("hamming_search")]


@pytest.mark.parametrize("r", HSEARCH, ids=_ids(HSEARCH))
def test_kat_hamming_search(eng, r):
    needle, hay = bytes.fromhex(r["a"]), bytes.fromhex(r["b"])
    if "k" in r:
        got = eng.hamming_search_simd_with_opts(needle, hay, r["k"], r["search_type"])
    else:
        got = eng.hamming_search(needle, hay)
    assert [tuple(m) for m in got] == [(m["start"], m["end"], m["k"]) for m in r["expect"]["matches"]]


def test_hamming_search_random(eng):
    rng = random.Random(31)
    for trial in range(12):
        alpha = rng.choice([2, 3, 20])
        nlen = rng.choice([1, 2, 5, 16, 33, 70])
        needle = bytes(1 + rng.randrange(alpha) for _ in range(nlen))
        hays = []
        for _ in range(100):
            h = bytearray(1 + rng.randrange(alpha) for _ in range(rng.randrange(0, 400)))
            if rng.random() < 0.5 and len(h) > nlen:
                p = rng.randrange(len(h) - nlen)
                h[p:p + nlen] = bytes((c if rng.random() < 0.9 else 1 + rng.randrange(alpha)) for c in needle)
            hays.append(bytes(h))
        hay, hoff = _pack(hays)
        for k in (0, 1, nlen // 2, nlen, nlen + 5):
            for st in (0, 1):
                got, goff = eng.hamming_search_batch(needle, hay, hoff, k, st)
                want, woff = orc.hamming_search_batch(needle, hay, hoff, k, st, threads=8)
                assert np.array_equal(goff, woff) and np.array_equal(got, want), (nlen, k, st)
    # empty needle -> no matches (src/hamming.rs:459-461); NUL byte in a haystack -> panic (src/lib.rs:237-243)
    hay, hoff = _pack([b"abc", b"defg"])
    got, goff = eng.hamming_search_batch(b"", hay, hoff, 1, 0)
    assert len(got) == 0 and list(goff) == [0, 0, 0]
    hay, hoff = _pack([b"abc", b"de\0g"])
    with pytest.raises(AssertionError):
        eng.hamming_search_batch(b"de", hay, hoff, 1, 0)


@pytest.mark.parametrize("nlen", [49 * 1024 + 3, 240 * 1024 + 5])
def test_hamming_search_large_needles(eng, nlen):
    """needles beyond the default 48 KB of dynamic shared memory (opt-in) and beyond what an SM has at all (the needle
    is then read from global memory); the reference takes any length (src/hamming.rs:440-520)"""
    rng = np.random.default_rng(nlen)
    needle = rng.integers(1, 255, size=nlen, dtype=np.uint8)
    a = rng.integers(1, 255, size=nlen + 300, dtype=np.uint8)
    a[100:100 + nlen] = needle
    a[100 + 17] ^= 1
    a[100 + nlen - 1] ^= 2
    b = rng.integers(1, 255, size=nlen - 1, dtype=np.uint8)   # shorter than the needle: no match
    c = needle.copy()                                          # exactly the needle
    hay, hoff = _pack([a.tobytes(), b.tobytes(), c.tobytes()])
    for k, st in ((2, 0), (1, 1), (5, 1)):
        got, goff = eng.hamming_search_batch(needle.tobytes(), hay, hoff, k, st)
        want, woff = orc.hamming_search_batch(needle.tobytes(), hay, hoff, k, st, threads=2)
        assert np.array_equal(goff, woff) and np.array_equal(got, want), (nlen, k, st)


# ---------------------------------------------------------------------------------------------------------------
# error behaviour of the boundary
def test_hamming_search_naive_accepts_nul_bytes(eng):
    """src/hamming.rs:96-146 has no NUL check (only the SIMD entry does, :463): haystacks with NUL bytes through the
    *_naive* names equal the scalar oracle's, through the SIMD names they are the reference's panic"""
    rng = random.Random(8)
    needle = bytes(rng.randrange(0, 3) for _ in range(6))
    hays = [bytes(rng.randrange(0, 3) for _ in range(rng.randrange(0, 300))) for _ in range(50)]
    hay, hoff = _pack(hays)
    for st in (0, 1):
        got, goff = eng.hamming_search_batch(needle, hay, hoff, 2, st, naive=True)
        for i, h in enumerate(hays):
            want = orc.hamming_search_naive_with_opts(needle, h, 2, st)
            assert [tuple(int(x) for x in m) for m in got[int(goff[i]):int(goff[i + 1])]] == want, (i, st)
    assert [tuple(m) for m in eng.hamming_search_naive_with_opts(needle, hays[3], 2, 0)] == \
        orc.hamming_search_naive_with_opts(needle, hays[3], 2, 0)
    with pytest.raises(AssertionError):
        eng.hamming_search_batch(needle, hay, hoff, 2, 0)


def test_hamming_length_mismatch_panics(eng):
    with pytest.raises(AssertionError):
        eng.hamming(b"abc", b"ab")  # reference: assert at src/hamming.rs:38
    a, ao = _pack([b"abcd", b"xy"])
    b, bo = _pack([b"abcd", b"xyz"])
    with pytest.raises(AssertionError):
        eng.hamming_batch(a, ao, b, bo)


def test_bad_costs_rejected(eng):
    from triple_accel_b200 import _ffi
    lib = _ffi.load()
    import ctypes as C
    out = C.c_uint32()
    rc = lib.ta_levenshtein_simd_k_with_opts(eng._h, b"a", 1, b"b", 1, 1, _ffi.ta_costs(0, 1, 0, 0), C.byref(out))
    assert rc == _ffi.TA_ERR_BAD_COSTS
    mp, op = C.POINTER(_ffi.ta_match)(), C.POINTER(C.c_uint64)()
    off = np.array([0, 1], np.uint64)
    rc = lib.ta_levenshtein_search_batch(eng._h, b"ab", 2, b"b", off.ctypes.data, 1, 1, 0, _ffi.ta_costs(2, 2, 0, 3), 0,
                                         C.byref(mp), C.byref(op))
    assert rc == _ffi.TA_ERR_BAD_COSTS  # check_search, src/levenshtein.rs:69


def test_empty_batch(eng):
    z8, z64 = np.zeros(0, np.uint8), np.zeros(1, np.uint64)
    assert len(eng.hamming_batch(z8, z64, z8, z64)) == 0
    assert len(eng.levenshtein_k_batch(z8, z64, z8, z64, 3)) == 0
    assert len(eng.levenshtein_exp_batch(z8, z64, z8, z64)) == 0


# ---------------------------------------------------------------------------------------------------------------
# randomised differential tests against the oracle
def _rand_strs(rng, n, lo, hi, alpha):
    return [bytes(rng.randrange(alpha) for _ in range(rng.randrange(lo, hi + 1))) for _ in range(n)]


def _mutate(rng, s, e, alpha, swap=True):
    s = bytearray(s)
    for _ in range(e):
        kind = rng.randrange(4 if swap else 3)
        if kind == 0 and s:
            s[rng.randrange(len(s))] = rng.randrange(alpha)
        elif kind == 1:
            s.insert(rng.randrange(len(s) + 1), rng.randrange(alpha))
        elif kind == 2 and s:
            del s[rng.randrange(len(s))]
        elif kind == 3 and len(s) > 1:
            p = rng.randrange(len(s) - 1)
            s[p], s[p + 1] = s[p + 1], s[p]
    return bytes(s)


@pytest.mark.parametrize("lens", [(0, 0), (1, 15), (16, 16), (64, 64), (0, 200), (1000, 1100), (4096, 4096)])
def test_hamming_random(eng, lens):
    rng = random.Random(hash(lens) & 0xFFFF)
    n = 3000 if lens[1] <= 256 else 200
    A = _rand_strs(rng, n, lens[0], lens[1], 4)
    B = [bytes((c if rng.random() < 0.7 else rng.randrange(256)) for c in s) for s in A]
    a, ao = _pack(A)
    b, bo = _pack(B)
    got = eng.hamming_batch(a, ao, b, bo)
    want = orc.hamming_batch(a, ao, b, bo)
    assert np.array_equal(got, want)


def test_hamming_misaligned_views(eng):
    """strings that start at odd addresses inside the byte buffers (CSR offsets need not be aligned)"""
    rng = random.Random(5)
    A = _rand_strs(rng, 500, 0, 90, 256)
    B = [bytes((c if rng.random() < 0.8 else rng.randrange(256)) for c in s) for s in A]
    a, ao = _pack(A)
    b, bo = _pack(B)
    a2 = np.concatenate([np.zeros(3, np.uint8), a])  # shift a by 3 bytes, b by 0
    got = eng.hamming_batch(a2, ao + np.uint64(3), b, bo)
    assert np.array_equal(got, orc.hamming_batch(a, ao, b, bo))


@pytest.mark.parametrize("costs", COST_MODELS, ids=[str(c) for c in COST_MODELS])
def test_lev_k_random_short(eng, costs):
    """short strings over small alphabets (many ties, transpositions and empty strings), every k regime"""
    rng = random.Random(sum(costs))
    A, B = [], []
    for _ in range(4000):
        alpha = rng.choice([2, 3, 4, 26])
        A.append(bytes(rng.randrange(alpha) for _ in range(rng.randrange(0, 24))))
        B.append(bytes(rng.randrange(alpha) for _ in range(rng.randrange(0, 24))))
    a, ao = _pack(A)
    b, bo = _pack(B)
    for k in (0, 1, 2, 3, 5, 8, 16, 30, 100, 0xFFFFFFFF):
        got = eng.levenshtein_k_batch(a, ao, b, bo, k, costs)
        want = orc.levenshtein_k_batch(a, ao, b, bo, k, costs)
        bad = np.nonzero(got != want)[0]
        assert len(bad) == 0, (k, costs, A[bad[0]], B[bad[0]], int(got[bad[0]]), int(want[bad[0]]))


@pytest.mark.parametrize("length,k", [(128, 8), (128, 16), (512, 16), (300, 30), (1024, 30), (100, 64), (77, 200),
                                      (2000, 5), (300, 45), (1024, 63), (600, 62), (200, 32),
                                      (130, 15), (131, 17), (200, 24), (96, 25), (47, 23)])
@pytest.mark.parametrize("costs", [(1, 1, 0, 0), (1, 1, 0, 1), (2, 1, 3, 0), (2, 2, 1, 3)], ids=str)
def test_lev_k_mutated(eng, length, k, costs):
    """the BASELINE shapes at reduced batch size: mutated pairs (within k) and unrelated pairs (None)"""
    rng = random.Random(length * 131 + k)
    n = 600 if length <= 512 else 150
    A = _rand_strs(rng, n, max(0, length - 3), length, 256)
    B = [(_mutate(rng, s, rng.randrange(0, k + 3), 256) if rng.random() < 0.8 else
          bytes(rng.randrange(256) for _ in range(len(s)))) for s in A]
    a, ao = _pack(A)
    b, bo = _pack(B)
    got = eng.levenshtein_k_batch(a, ao, b, bo, k, costs)
    want = orc.levenshtein_k_batch(a, ao, b, bo, k, costs, threads=8)
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, (len(bad), A[bad[0]], B[bad[0]], int(got[bad[0]]), int(want[bad[0]]))
    # swapped argument order gives the same answers (the reference swaps so that a is the shorter string)
    got2 = eng.levenshtein_k_batch(b, bo, a, ao, k, costs)
    assert np.array_equal(got2, want)


@pytest.mark.parametrize("costs", COST_MODELS, ids=[str(c) for c in COST_MODELS])
def test_lev_diag16_large_batch(eng, costs):
    """a batch large enough for the dispatcher's thread-per-pair u16 kernel (lev_diag16.cu; >= 16384 pairs, band <= 32
    diagonals): ragged lengths incl. empty strings, every misalignment, small alphabets (ties, transpositions), the
    bands of every register count (2, 3, 4, 6, 8 packed registers per anti-diagonal), None and within-k pairs"""
    from triple_accel_b200 import synth
    parts = [synth.edited_pairs(6000, 0, 70, 6, seed=sum(costs), allow_swap=True, alphabet=4),
             synth.edited_pairs(6000, 90, 140, 12, seed=sum(costs) + 1, allow_swap=True),
             synth.edited_pairs(5000, 1, 40, 3, seed=sum(costs) + 2, allow_swap=True, alphabet=2)]
    a = np.concatenate([p[0] for p in parts])
    b = np.concatenate([p[2] for p in parts])
    ao = np.concatenate([[0]] + [p[1][1:] + sum(len(q[0]) for q in parts[:i]) for i, p in enumerate(parts)]).astype(np.uint64)
    bo = np.concatenate([[0]] + [p[3][1:] + sum(len(q[2]) for q in parts[:i]) for i, p in enumerate(parts)]).astype(np.uint64)
    mism, gap, sgap, _ = costs
    for k in (0, 2, 5, 9, 14, 21, 30):
        want = orc.levenshtein_k_batch(a, ao, b, bo, k, costs, threads=8)
        got = eng.levenshtein_k_batch(a, ao, b, bo, k, costs)
        bad = np.nonzero(got != want)[0]
        assert len(bad) == 0, (k, costs, len(bad), int(bad[0]), int(got[bad[0]]), int(want[bad[0]]),
                               bytes(a[int(ao[bad[0]]):int(ao[bad[0] + 1])]), bytes(b[int(bo[bad[0]]):int(bo[bad[0] + 1])]))
    assert np.array_equal(eng.levenshtein_k_batch(b, bo, a, ao, 9, costs), orc.levenshtein_k_batch(a, ao, b, bo, 9, costs, threads=8))


@pytest.mark.parametrize("costs", [(1, 1, 0, 0), (1, 1, 0, 1)], ids=str)
@pytest.mark.parametrize("alpha", [2, 4, 256])
def test_lev_fr_long_strings(eng, costs, alpha):
    """long strings, small k: the dispatcher's diagonal-extension kernel (lev_fr.cu).  Small alphabets and periodic
    strings make many diagonals slide far (every lane of an octet queues a cooperative slide); ragged lengths and
    offsets exercise every 16-byte re-alignment of the two streams; swapped arguments, None and length-difference
    exits included."""
    rng = random.Random(alpha * 7 + costs[3])
    A, B = [], []
    for i in range(260):
        la = rng.choice((1024, 1500, 2048, 4096)) + rng.randrange(0, 40)
        if i % 5 == 4:  # periodic: every diagonal that is a multiple of the period matches for ever
            unit = bytes(rng.randrange(alpha) for _ in range(rng.choice((1, 2, 3, 7, 16))))
            s = (unit * (la // len(unit) + 1))[:la]
        else:
            s = bytes(rng.randrange(alpha) for _ in range(la))
        r = rng.random()
        if r < 0.75:
            t = _mutate(rng, s, rng.randrange(0, 20), alpha)
        elif r < 0.85:
            t = s
        else:
            t = bytes(rng.randrange(alpha) for _ in range(la + rng.randrange(-5, 6)))
        A.append(s)
        B.append(t)
    a, ao = _pack(A)
    b, bo = _pack(B)
    for k in (0, 3, 8, 16):
        want = orc.levenshtein_k_batch(a, ao, b, bo, k, costs, threads=8)
        got = eng.levenshtein_k_batch(a, ao, b, bo, k, costs)
        bad = np.nonzero(got != want)[0]
        assert len(bad) == 0, (k, len(bad), int(bad[0]), int(got[bad[0]]), int(want[bad[0]]))
        assert np.array_equal(eng.levenshtein_k_batch(b, bo, a, ao, k, costs), want)
    assert np.array_equal(eng.levenshtein_exp_batch(a, ao, b, bo, costs),
                          orc.levenshtein_exp_batch(a, ao, b, bo, costs, threads=8))


@pytest.mark.parametrize("costs", [(1, 1, 0, 0), (1, 1, 0, 1), (1, 1, 2, 0), (3, 2, 0, 2)], ids=str)
def test_lev_full_matrix(eng, costs):
    """levenshtein() / rdamerau(): k = u32::MAX, the band is the whole matrix"""
    rng = random.Random(99)
    A = _rand_strs(rng, 300, 0, 200, 5)
    B = _rand_strs(rng, 300, 0, 200, 5)
    a, ao = _pack(A)
    b, bo = _pack(B)
    got = eng.levenshtein_k_batch(a, ao, b, bo, 0xFFFFFFFF, costs)
    want = orc.levenshtein_k_batch(a, ao, b, bo, 0xFFFFFFFF, costs, threads=8)
    assert np.array_equal(got, want)
    assert not (got == NONE).any()


@pytest.mark.parametrize("costs", [(1, 1, 0, 0), (1, 1, 0, 1), (2, 1, 2, 0)], ids=str)
def test_lev_exp(eng, costs):
    rng = random.Random(17)
    A = _rand_strs(rng, 400, 0, 300, 256)
    B = []
    for s in A:
        r = rng.random()
        if r < 0.5:
            B.append(_mutate(rng, s, rng.randrange(0, 6), 256))
        elif r < 0.8:
            B.append(_mutate(rng, s, rng.randrange(20, 90), 256))  # needs k = 60, 120 rounds
        else:
            B.append(bytes(rng.randrange(256) for _ in range(rng.randrange(0, 300))))
    a, ao = _pack(A)
    b, bo = _pack(B)
    got = eng.levenshtein_exp_batch(a, ao, b, bo, costs)
    want = orc.levenshtein_exp_batch(a, ao, b, bo, costs, threads=8)
    assert np.array_equal(got, want)


def test_nul_bytes_and_high_bytes(eng):
    """the reference guards its zero-padded SIMD windows with NUL-byte cases (tests/basic_tests.rs:503-537)"""
    rng = random.Random(3)
    A = [bytes(rng.choice([0, 0, 255, 1]) for _ in range(rng.randrange(0, 40))) for _ in range(1500)]
    B = [bytes(rng.choice([0, 0, 255, 1]) for _ in range(rng.randrange(0, 40))) for _ in range(1500)]
    a, ao = _pack(A)
    b, bo = _pack(B)
    for costs in ((1, 1, 0, 0), (1, 1, 0, 1)):
        for k in (2, 7, 0xFFFFFFFF):
            assert np.array_equal(eng.levenshtein_k_batch(a, ao, b, bo, k, costs),
                                  orc.levenshtein_k_batch(a, ao, b, bo, k, costs))


SEARCH_MODELS = [(1, 1, 0, 0), (1, 1, 0, 1), (1, 1, 2, 0), (2, 1, 2, 0), (3, 1, 0, 0), (2, 2, 1, 3), (2, 2, 0, 2)]


@pytest.mark.parametrize("costs", SEARCH_MODELS, ids=[str(c) for c in SEARCH_MODELS])
@pytest.mark.parametrize("anchored", [False, True])
def test_search_random(eng, costs, anchored):
    rng = random.Random(sum(costs) + anchored)
    for trial in range(6):
        alpha = rng.choice([2, 3, 4, 20])
        nlen = rng.choice([1, 2, 3, 5, 8, 13, 32, 40])
        needle = bytes(rng.randrange(alpha) for _ in range(nlen))
        hays = []
        for _ in range(120):
            h = bytearray(rng.randrange(alpha) for _ in range(rng.randrange(0, 120)))
            if rng.random() < 0.5 and len(h) > nlen:
                p = rng.randrange(len(h) - nlen)
                h[p:p + nlen] = _mutate(rng, needle, rng.randrange(0, 3), alpha)
            hays.append(bytes(h))
        hay, hoff = _pack(hays)
        for k in (0, 1, 2, nlen // 2 + 1, 3 * nlen):
            for st in (0, 1):
                got, goff = eng.levenshtein_search_batch(needle, hay, hoff, k, st, costs, anchored)
                want, woff = orc.levenshtein_search_batch(needle, hay, hoff, k, st, costs, anchored, threads=8)
                assert np.array_equal(goff, woff), (needle, k, st, costs, anchored)
                assert np.array_equal(got, want), (needle, k, st, costs, anchored)


def test_search_planted_long_haystacks(eng):
    """cfg 4 shape at reduced size: needle 32, haystacks of 4096 bytes 1..255, k = 3, planted hits"""
    from triple_accel_b200 import synth
    needle, hay, hoff = synth.needle_haystacks(600, 4096, 32, plant_frac=0.05, max_edits=3, seed=7)
    for st in (0, 1):
        got, goff = eng.levenshtein_search_batch(needle, hay, hoff, 3, st)
        want, woff = orc.levenshtein_search_batch(needle, hay, hoff, 3, st, threads=8)
        assert np.array_equal(goff, woff) and np.array_equal(got, want)
        assert goff[-1] >= 20  # the planted needles are found


def test_device_resident_entry_points(eng):
    import torch
    from triple_accel_b200 import synth
    a, ao, b, bo = synth.mutated_pairs(20000, 128, 8, seed=5)
    dev = torch.device("cuda", eng.device)
    ta_, tao, tb, tbo = (torch.from_numpy(x.view(np.int64) if x.dtype == np.uint64 else x).to(dev) for x in (a, ao, b, bo))
    out = torch.empty(20000, dtype=torch.int32, device=dev)
    eng.levenshtein_k_batch_dev(ta_, tao, tb, tbo, 8, (1, 1, 0, 0), 136, out)
    eng.dev_status()
    want = orc.levenshtein_k_batch(a, ao, b, bo, 8, threads=8)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want)
    # hamming on the equal-length prefix pairs
    a2, ao2, b2, bo2 = synth.hamming_pairs(10000, 64, seed=9)
    t = [torch.from_numpy(x.view(np.int64) if x.dtype == np.uint64 else x).to(dev) for x in (a2, ao2, b2, bo2)]
    out2 = torch.empty(10000, dtype=torch.int32, device=dev)
    eng.hamming_batch_dev(t[0], t[1], t[2], t[3], out2)
    eng.dev_status()
    assert np.array_equal(out2.cpu().numpy().view(np.uint32), orc.hamming_batch(a2, ao2, b2, bo2))
    for mean_len in (1, 16, 64, 300, 5000):  # ta_hamming_batch_dev_len: lanes per pair follow the caller's mean length
        out2.fill_(9)
        eng.hamming_batch_dev(t[0], t[1], t[2], t[3], out2, mean_len=mean_len)
        eng.dev_status()
        assert np.array_equal(out2.cpu().numpy().view(np.uint32), orc.hamming_batch(a2, ao2, b2, bo2)), mean_len
    # a length mismatch on the device path is reported by ta_dev_status
    eng.hamming_batch_dev(t[0], t[1], tb, tbo, out2)
    with pytest.raises(AssertionError):
        eng.dev_status()


def test_exp_and_search_device_entry_points(eng):
    import torch
    from triple_accel_b200 import synth
    dev = torch.device("cuda", eng.device)

    def to_dev(x):
        return torch.from_numpy(x.view(np.int64) if x.dtype == np.uint64 else x).to(dev)

    a, ao, b, bo = synth.mutated_pairs(3000, 300, 50, seed=11)
    out = torch.empty(3000, dtype=torch.int32, device=dev)
    eng.levenshtein_exp_batch_dev(to_dev(a), to_dev(ao), to_dev(b), to_dev(bo), (1, 1, 0, 0), 400, out)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), orc.levenshtein_exp_batch(a, ao, b, bo, threads=8))
    needle, hay, hoff = synth.needle_haystacks(400, 2048, 32, plant_frac=0.1, seed=12)
    for st in (0, 1):
        got, goff = eng.levenshtein_search_batch_dev(needle, to_dev(hay), to_dev(hoff), 2048, 3, st)
        want, woff = orc.levenshtein_search_batch(needle, hay, hoff, 3, st, threads=8)
        assert np.array_equal(goff, woff) and np.array_equal(got, want)


def test_search_segment_warmup_is_exact(eng):
    """With the pre-filter, the exact (cost, length) DP is re-run per flagged 512-byte segment after a warm-up of
    2N + 2 bytes instead of from the start of the haystack.  Tiny alphabets make length ties the common case, so
    any dependence of a match's `start` on bytes before the warm-up window would show up here."""
    rng = random.Random(20261017)
    for trial in range(10):
        alpha = rng.choice([2, 2, 3, 4])
        nlen = rng.choice([3, 5, 8, 13, 21, 32, 40, 64])
        needle = bytes(rng.randrange(alpha) for _ in range(nlen))
        hays = []
        for _ in range(40):
            h = bytearray(rng.randrange(alpha) for _ in range(rng.randrange(1500, 3200)))
            for _ in range(rng.randrange(0, 6)):
                p = rng.randrange(len(h) - nlen)
                h[p:p + nlen] = _mutate(rng, needle, rng.randrange(0, 4), alpha)[:nlen]
            hays.append(bytes(h))
        hay, hoff = _pack(hays)
        for costs in ((1, 1, 0, 0), (1, 1, 0, 1)):
            for k in (0, 1, max(1, nlen // 5), nlen // 2):
                if k >= nlen:
                    continue
                for st in (0, 1):
                    got, goff = eng.levenshtein_search_batch(needle, hay, hoff, k, st, costs)
                    want, woff = orc.levenshtein_search_batch(needle, hay, hoff, k, st, costs, threads=8)
                    assert np.array_equal(goff, woff) and np.array_equal(got, want), (alpha, nlen, costs, k, st)


def test_search_filter_long_needles_and_transpositions(eng):
    """the bit-parallel pre-filter (needle <= 64, unit costs, with and without transpositions) must never drop a
    haystack that has a match: compare the full match lists with the oracle"""
    rng = random.Random(77)
    for nlen in (7, 31, 32, 33, 48, 64, 65):
        for costs in ((1, 1, 0, 0), (1, 1, 0, 1)):
            alpha = 4
            needle = bytes(rng.randrange(alpha) for _ in range(nlen))
            hays = []
            for _ in range(150):
                h = bytearray(rng.randrange(alpha) for _ in range(rng.randrange(0, 1500)))
                if rng.random() < 0.4 and len(h) > nlen + 2:
                    p = rng.randrange(len(h) - nlen)
                    h[p:p + nlen] = _mutate(rng, needle, rng.randrange(0, 5), alpha)[:nlen]
                hays.append(bytes(h))
            hay, hoff = _pack(hays)
            for k in (0, 2, nlen // 4, nlen - 1):
                got, goff = eng.levenshtein_search_batch(needle, hay, hoff, k, 0, costs)
                want, woff = orc.levenshtein_search_batch(needle, hay, hoff, k, 0, costs, threads=8)
                assert np.array_equal(goff, woff) and np.array_equal(got, want), (nlen, costs, k)


def test_search_qgram_filter(eng):
    """the aligned-word (4-gram) pre-filter: needles of up to 64 bytes whose k + 1 (2k + 1) pieces are >= 7 bytes.
    The dispatcher takes it from 64 MB (8 MB for needles > 32) of haystacks per call -- test_search_qgram_big_batch runs
    it at that size; here the forced-variant run (TA_SEARCH_FILTER=qgram) is what exercises it, the plain run pins the
    scanning filters on the same inputs.  Alphabets of 2 / 4
    symbols make every word a candidate (the queue, the whole-piece compare and the verification do the work; big
    batches overflow the queue and take the shift-and fallback), 256 symbols is the case it is built for; haystacks
    are ragged, start at every alignment, and include empty and shorter-than-a-word ones."""
    rng = random.Random(1717)
    for nlen, k, costs in ((14, 1, (1, 1, 0, 0)), (21, 2, (1, 1, 0, 0)), (32, 3, (1, 1, 0, 0)), (28, 3, (1, 1, 0, 0)),
                           (32, 1, (1, 1, 0, 1)), (21, 1, (1, 1, 0, 1)), (32, 6, (2, 2, 0, 0)), (30, 0, (1, 1, 0, 0)),
                           (64, 6, (1, 1, 0, 0)), (48, 5, (1, 1, 0, 0)), (33, 3, (1, 1, 0, 0)), (64, 3, (1, 1, 0, 1)),
                           (57, 0, (1, 1, 0, 1))):
        for alpha in (2, 4, 256):
            needle = bytes(rng.randrange(alpha) for _ in range(nlen))
            hays = [b"", b"ab", needle, needle[1:], needle + needle]
            for _ in range(300):
                h = bytearray(rng.randrange(alpha) for _ in range(rng.randrange(0, 700)))
                for _ in range(rng.randrange(0, 3)):
                    if len(h) > nlen + 2:
                        pos = rng.randrange(len(h) - nlen)
                        h[pos:pos + nlen] = _mutate(rng, needle, rng.randrange(0, 5), alpha)[:nlen]
                hays.append(bytes(h))
            hay, hoff = _pack(hays)
            for st in (0, 1):
                got, goff = eng.levenshtein_search_batch(needle, hay, hoff, k, st, costs)
                want, woff = orc.levenshtein_search_batch(needle, hay, hoff, k, st, costs, threads=8)
                assert np.array_equal(goff, woff) and np.array_equal(got, want), (nlen, k, costs, alpha, st)


@pytest.mark.parametrize("nlen,k,alpha,plant", [(32, 3, 256, 0.05), (64, 6, 256, 0.05), (32, 3, 4, 0.05), (32, 3, 256, 1.0)])
def test_search_qgram_big_batch(eng, nlen, k, alpha, plant):
    """70 MB of haystacks: the default dispatch takes the q-gram scan (alphabet 4: its queue overflows and the fallback
    kernel runs; plant = 1.0: a match in EVERY haystack must still fit the queue); the oracle checks a slice, the rest through the planted needles every haystack of the slice pattern
    must report"""
    from triple_accel_b200 import synth
    n, hlen = 18000, 4096
    needle = np.random.default_rng(nlen).integers(1, 256, size=nlen, dtype=np.uint8)
    hay, hoff = synth.planted_haystacks(n, hlen, needle, plant_frac=plant, max_edits=k, seed=nlen + alpha)
    if alpha != 256:  # the same map on both sides keeps the planted copies planted
        hay = (hay % alpha).astype(np.uint8)
        needle = (needle % alpha).astype(np.uint8)
    needle = needle.tobytes()
    got, goff = eng.levenshtein_search_batch(needle, hay, hoff, k, 1)
    chk = 600
    want, woff = orc.levenshtein_search_batch(needle, np.asarray(hay)[: chk * hlen], np.asarray(hoff)[: chk + 1], k, 1, threads=8)
    assert np.array_equal(goff[: chk + 1], woff) and np.array_equal(got[: int(woff[-1])], want)
    assert int(goff[-1]) >= int(woff[-1])


@pytest.mark.parametrize("nlen", [257, 301, 453, 1000])
@pytest.mark.parametrize("costs", [(1, 1, 0, 0), (1, 1, 1, 1)], ids=str)
def test_search_long_needles(eng, nlen, costs):
    """needles beyond both exact kernels' on-chip limits (256 rows for the warp kernel, ~450 / ~300 rows of shared
    memory for the thread kernel): DP rows in the global-memory workspace.  The reference takes any needle
    (src/levenshtein.rs:2034-2151)."""
    rng = random.Random(nlen + costs[3])
    needle = bytes(rng.randrange(1, 5) for _ in range(nlen))
    hays = []
    for i in range(40):
        h = bytearray(rng.randrange(1, 5) for _ in range(rng.randrange(0, 2500)))
        if i % 3 == 0:
            p = rng.randrange(0, len(h) + 1)
            h[p:p] = _mutate(rng, needle, rng.randrange(0, 30), 5).replace(b"\0", b"\1")
        hays.append(bytes(h))
    hay, hoff = _pack(hays)
    for k, st, anchored in ((25, 0, False), (25, 1, False), (40, 1, True)):
        got, goff = eng.levenshtein_search_batch(needle, hay, hoff, k, st, costs, anchored)
        want, woff = orc.levenshtein_search_batch(needle, hay, hoff, k, st, costs, anchored, threads=8)
        assert np.array_equal(goff, woff) and np.array_equal(got, want), (nlen, k, st, anchored)


@pytest.mark.parametrize("costs", [(1, 1, 0, 0), (1, 1, 0, 1), (2, 1, 3, 0), (2, 2, 1, 3)], ids=str)
def test_lev_wide_band_block_kernel(eng, costs):
    """bands wider than 1024 diagonals go to the block-per-pair kernel: levenshtein()/rdamerau() on long strings
    (k = u32::MAX) and large explicit k"""
    rng = random.Random(4242)
    A = _rand_strs(rng, 12, 600, 1500, 4) + [b"", b"x" * 1200]
    B = [(_mutate(rng, s, rng.randrange(0, 700), 4) if rng.random() < 0.6 else
          bytes(rng.randrange(4) for _ in range(rng.randrange(500, 1500)))) for s in A[:12]] + [b"y" * 1300, b""]
    a, ao = _pack(A)
    b, bo = _pack(B)
    for k in (0xFFFFFFFF, 1100):
        got = eng.levenshtein_k_batch(a, ao, b, bo, k, costs)
        want = orc.levenshtein_k_batch(a, ao, b, bo, k, costs, threads=8)
        assert np.array_equal(got, want), (k, costs, got, want)
    assert np.array_equal(eng.levenshtein_exp_batch(a, ao, b, bo, costs),
                          orc.levenshtein_exp_batch(a, ao, b, bo, costs, threads=8))


def test_lev_very_wide_band_global_workspace(eng):
    """anti-diagonals that do not fit shared memory use the per-block global workspace"""
    rng = random.Random(99)
    A = [bytes(rng.randrange(3) for _ in range(6100)) for _ in range(2)]
    B = [bytes(rng.randrange(3) for _ in range(6000)), _mutate(rng, A[1], 300, 3)]
    a, ao = _pack(A)
    b, bo = _pack(B)
    got = eng.levenshtein_k_batch(a, ao, b, bo, 0xFFFFFFFF, (1, 1, 0, 1))
    want = orc.levenshtein_k_batch(a, ao, b, bo, 0xFFFFFFFF, (1, 1, 0, 1), threads=2)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("costs", COST_MODELS, ids=[str(c) for c in COST_MODELS])
def test_traceback_random(eng, costs):
    """trace_on = true: the run-length encoded edits must equal the scalar routine's, tie-breaks included
    (src/levenshtein.rs:493-606); small alphabets make ties the common case"""
    rng = random.Random(1000 + sum(costs))
    A, B = [], []
    for _ in range(1500):
        alpha = rng.choice([2, 3, 4, 26])
        s = bytes(rng.randrange(alpha) for _ in range(rng.randrange(0, 40)))
        A.append(s)
        B.append(_mutate(rng, s, rng.randrange(0, 8), alpha) if rng.random() < 0.7 else
                 bytes(rng.randrange(alpha) for _ in range(rng.randrange(0, 40))))
    a, ao = _pack(A)
    b, bo = _pack(B)
    for k in (0, 3, 9, 40, 0xFFFFFFFF):
        dist, edits, eoff = eng.levenshtein_k_trace_batch(a, ao, b, bo, k, costs)
        assert np.array_equal(dist, orc.levenshtein_k_batch(a, ao, b, bo, k, costs))
        for i in range(0, len(A), 3):
            want = orc.levenshtein_naive_k_with_opts(A[i], B[i], k, True, costs)
            got = [tuple(int(x) for x in e) for e in edits[int(eoff[i]):int(eoff[i + 1])]]
            if want is None:
                assert dist[i] == NONE and got == []
            else:
                assert got == [tuple(e) for e in want[1]], (A[i], B[i], k, costs, got, want)
    # exponential search with traceback (src/levenshtein.rs:1480-1494)
    dist, edits, eoff = eng.levenshtein_exp_trace_batch(a, ao, b, bo, costs)
    assert np.array_equal(dist, orc.levenshtein_exp_batch(a, ao, b, bo, costs))
    for i in range(0, len(A), 7):
        want = orc.levenshtein_naive_k_with_opts(A[i], B[i], 0xFFFFFFFF, True, costs)
        got = [tuple(int(x) for x in e) for e in edits[int(eoff[i]):int(eoff[i + 1])]]
        assert got == [tuple(e) for e in want[1]], (A[i], B[i], costs)


def test_traceback_longer_strings(eng):
    rng = random.Random(5150)
    A = _rand_strs(rng, 120, 100, 400, 4)
    B = [_mutate(rng, s, rng.randrange(0, 60), 4) for s in A]
    a, ao = _pack(A)
    b, bo = _pack(B)
    for costs, k in (((1, 1, 0, 0), 30), ((1, 1, 0, 1), 70), ((2, 1, 3, 0), 200), ((1, 1, 0, 0), 0xFFFFFFFF)):
        dist, edits, eoff = eng.levenshtein_k_trace_batch(a, ao, b, bo, k, costs)
        for i in range(len(A)):
            want = orc.levenshtein_naive_k_with_opts(A[i], B[i], k, True, costs)
            got = [tuple(int(x) for x in e) for e in edits[int(eoff[i]):int(eoff[i + 1])]]
            if want is None:
                assert dist[i] == NONE and got == []
            else:
                assert dist[i] == want[0] and got == [tuple(e) for e in want[1]], (i, costs, k)


def test_traceback_long_strings_unbounded_k(eng):
    """levenshtein() / rdamerau()-style callers ask for the trace with k = u32::MAX: the band of the traceback must
    follow the pairs' distances, not k or the string lengths (strings of 600..3000 bytes, a handful of edits; one
    unrelated pair of short strings whose distance is large)."""
    rng = random.Random(77)
    A = _rand_strs(rng, 40, 600, 3000, 256)
    B = [_mutate(rng, s, rng.randrange(0, 9), 256) for s in A]
    A.append(bytes(rng.randrange(256) for _ in range(150)))
    B.append(bytes(rng.randrange(256) for _ in range(140)))
    a, ao = _pack(A)
    b, bo = _pack(B)
    for costs in ((1, 1, 0, 0), (1, 1, 0, 1), (2, 1, 2, 0)):
        for exp in (False, True):
            if exp:
                dist, edits, eoff = eng.levenshtein_exp_trace_batch(a, ao, b, bo, costs)
            else:
                dist, edits, eoff = eng.levenshtein_k_trace_batch(a, ao, b, bo, 0xFFFFFFFF, costs)
            for i in range(len(A)):
                wd, we = orc.levenshtein_naive_k_with_opts(A[i], B[i], 0xFFFFFFFF, True, costs)  # the SIMD entry's tie order
                got = [tuple(int(x) for x in e) for e in edits[int(eoff[i]):int(eoff[i + 1])]]
                assert dist[i] == wd and got == [tuple(e) for e in we], (i, costs, exp)


@pytest.mark.parametrize("n", [1, 2047, 2048, 2049, 70001])
def test_lev_duo_ragged_tiles(eng, n):
    """two pairs per thread behind the tile-local length ordering (lev_bitpar_duo_tiled_kernel): ragged batches with
    empty strings, batch sizes around the 2048-pair tile, one-class tiles (identity order) and many-class tiles, k <= 8"""
    from triple_accel_b200 import synth
    for lo, hi, seed in ((0, 200, 1), (128, 128, 2), (96, 160, 3), (1, 17, 4)):
        a, ao, b, bo = synth.edited_pairs(n, lo, hi, 8, seed=seed + n, allow_swap=False)
        for k, costs in ((8, (1, 1, 0, 0)), (3, (1, 1, 0, 0)), (16, (1, 1, 0, 0)), (8, (1, 1, 0, 1)), (22, (1, 1, 0, 1))):
            got = eng.levenshtein_k_batch(a, ao, b, bo, k, costs)  # k > 8 / transpositions: the one-pair-per-thread kernels
            want = orc.levenshtein_k_batch(a, ao, b, bo, k, costs, threads=8)
            assert np.array_equal(got, want), (n, lo, hi, k, costs)


def test_length_hint_changes_the_kernel_not_the_result(eng):
    """ta_set_length_hint: device-resident calls take the tile-ordered kernel only when told that lengths vary; the
    host-buffer calls decide from the offsets.  Same distances either way."""
    import torch
    from triple_accel_b200 import synth
    a, ao, b, bo = synth.edited_pairs(30000, 40, 200, 8, seed=99, allow_swap=False)
    want = orc.levenshtein_k_batch(a, ao, b, bo, 8, threads=8)
    dev = torch.device("cuda:0")

    def td(x):
        return torch.from_numpy(x.view(np.int64) if x.dtype == np.uint64 else x).to(dev)
    d_a, d_ao, d_b, d_bo = td(a), td(ao), td(b), td(bo)
    d_out = torch.empty(len(ao) - 1, dtype=torch.int32, device=dev)
    try:
        for hint in (True, False, None):
            eng.set_length_hint(hint)
            d_out.fill_(7)
            eng.levenshtein_k_batch_dev(d_a, d_ao, d_b, d_bo, 8, (1, 1, 0, 0), 200, d_out)
            torch.cuda.synchronize()
            assert np.array_equal(d_out.cpu().numpy().view(np.uint32), want), hint
            assert np.array_equal(eng.levenshtein_k_batch(a, ao, b, bo, 8), want), hint
    finally:
        eng.set_length_hint(None)


def test_full_size_properties(eng):
    """BASELINE.json sizes (1 M pairs, len 128, k = 8 and 16; 1 M x len 512 Damerau is covered by the bench's parity
    check) through size-independent properties: symmetry, identity, the edit budget as an upper bound, monotonicity
    in k, agreement between the two bit-parallel kernels' dispatch and the exponential search, plus an oracle check
    of a 30 k sample."""
    from triple_accel_b200 import synth
    n = 1_000_000
    a, ao, b, bo = synth.mutated_pairs(n, 128, 8, seed=77)
    d8 = eng.levenshtein_k_batch(a, ao, b, bo, 8)
    assert not (d8 == NONE).any() and d8.max() <= 8          # b = a after <= 8 edits
    assert np.array_equal(d8, eng.levenshtein_k_batch(b, bo, a, ao, 8))  # symmetry
    assert not eng.levenshtein_k_batch(a, ao, a, ao, 0).any()             # identity at k = 0
    d16 = eng.levenshtein_k_batch(a, ao, b, bo, 16)
    assert np.array_equal(d16, d8)                                        # monotone in k: same value once within k
    d3 = eng.levenshtein_k_batch(a, ao, b, bo, 3)
    assert np.array_equal(d3, np.where(d8 <= 3, d8, NONE).astype(np.uint32))
    assert np.array_equal(eng.levenshtein_exp_batch(a, ao, b, bo), d8)
    dd = eng.levenshtein_k_batch(a, ao, b, bo, 8, (1, 1, 0, 1))
    assert (dd <= d8).all()                                               # transpositions can only help
    sel = slice(123_000, 153_000)
    sa, sao = a[int(ao[sel.start]):int(ao[sel.stop])], ao[sel.start:sel.stop + 1] - ao[sel.start]
    sb, sbo = b[int(bo[sel.start]):int(bo[sel.stop])], bo[sel.start:sel.stop + 1] - bo[sel.start]
    assert np.array_equal(d8[sel], orc.levenshtein_k_batch(sa, sao, sb, sbo, 8, threads=8))
    assert np.array_equal(dd[sel], orc.levenshtein_k_batch(sa, sao, sb, sbo, 8, (1, 1, 0, 1), threads=8))
    # Hamming at full size: equal-length pairs, hamming >= levenshtein, hamming(a, a) = 0
    ha, hao, hb, hbo = synth.hamming_pairs(n, 128, seed=78)
    h = eng.hamming_batch(ha, hao, hb, hbo)
    assert h.max() <= 6 and not eng.hamming_batch(ha, hao, ha, hao).any()
    assert (eng.levenshtein_k_batch(ha, hao, hb, hbo, 8) <= h).all()


def test_full_size_search_properties(eng):
    """cfg 4 at full size (100 k haystacks x 4096 B): All vs Best consistency and an oracle check on the haystacks
    that report matches plus a random sample of silent ones."""
    from triple_accel_b200 import synth
    needle, hay, hoff = synth.needle_haystacks(100_000, 4096, 32, plant_frac=0.01, max_edits=3, seed=79)
    m_all, o_all = eng.levenshtein_search_batch(needle, hay, hoff, 3, 0)
    m_best, o_best = eng.levenshtein_search_batch(needle, hay, hoff, 3, 1)
    cnt_all, cnt_best = np.diff(o_all.astype(np.int64)), np.diff(o_best.astype(np.int64))
    assert ((cnt_all > 0) == (cnt_best > 0)).all() and (cnt_best <= cnt_all).all()
    assert 900 <= (cnt_all > 0).sum() <= 1100   # ~1 % planted
    assert (m_all[:, 2] <= 3).all() and (m_all[:, 0] <= m_all[:, 1]).all()
    hit = np.nonzero(cnt_all > 0)[0]
    rng = np.random.default_rng(1)
    chk = np.concatenate([hit, rng.choice(100_000, 300, replace=False)])
    for h in chk[:1500]:
        hs = bytes(hay[int(hoff[h]):int(hoff[h + 1])])
        for st, (mm, oo) in ((0, (m_all, o_all)), (1, (m_best, o_best))):
            want = orc.levenshtein_search_naive_with_opts(bytes(needle), hs, 3, st)
            got = [tuple(int(x) for x in r) for r in mm[int(oo[h]):int(oo[h + 1])]]
            assert got == want, (h, st)


def test_cpp_header_mirror(tmp_path):
    """include/triple_accel.hpp (the C++ host mirror of the crate's names) compiled against the shared library"""
    import shutil
    import subprocess
    if not shutil.which("g++"):
        pytest.skip("no g++ on this box")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "hpp_kat")
    libdir = os.path.join(root, "triple_accel_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "hpp_kat.cpp"), "-L", libdir, "-ltriple_accel_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "hpp ok" in r.stdout, r.stdout + r.stderr


LEV_TESTS = "test_lev_k_mutated or test_lev_k_random_short or test_nul_bytes or test_lev_exp or test_lev_duo_ragged_tiles"
FR_TESTS = LEV_TESTS + " or test_lev_fr_long_strings or test_lev_full_matrix"
SEARCH_TESTS = ("test_search_random or test_search_planted or test_kat_search or test_search_qgram_filter or "
                "test_search_filter_long_needles_and_transpositions or test_search_segment_warmup")


@pytest.mark.parametrize("env,select", [
    ({"TA_FORCE_BAND": "1"}, LEV_TESTS),
    ({"TA_FR": "1"}, FR_TESTS),
    ({"TA_DIAG16": "1"}, LEV_TESTS + " or test_lev_diag16_large_batch"),
    ({"TA_DIAG16": "1", "TA_FORCE_BAND": "1"}, LEV_TESTS),
    ({"TA_DIAG16": "1", "TA_DIAG16_STAGES": "1"}, "test_lev_diag16_large_batch or test_lev_k_mutated"),
    ({"TA_DIAG16": "0"}, "test_lev_diag16_large_batch"),
    ({"TA_FR": "0"}, "test_lev_fr_long_strings"),
    ({"TA_BITPAR": "simd"}, LEV_TESTS),
    ({"TA_BITPAR": "tab"}, LEV_TESTS),
    ({"TA_BITPAR": "tab", "TA_BITPAR_BITS": "32"}, LEV_TESTS),
    ({"TA_BLK_DUO": "0", "TA_EXP_FIRST_K": "30"}, LEV_TESTS),
    ({"TA_DUO_TILED": "0"}, LEV_TESTS + " or test_full_size"),
    ({"TA_DUO_TILED": "1"}, LEV_TESTS + " or test_full_size"),
    ({"TA_BLK_TILED": "1"}, LEV_TESTS + " or test_full_size"),
    ({"TA_BLK_TILED": "0", "TA_DUO_TILED": "0"}, LEV_TESTS),
    ({"TA_LEN_BUCKETS": "1"}, LEV_TESTS + " or test_full_size"),
    ({"TA_BLK_PLANES": "0"}, LEV_TESTS),
    ({"TA_BLK_PLANES": "1", "TA_BLK_C": "8"}, LEV_TESTS),
    ({"TA_BLK_PLANES": "0", "TA_BLK_C": "8", "TA_BITPAR_THREADS": "64"}, LEV_TESTS),
    ({"TA_BITPAR": "tab", "TA_BITPAR_BITS": "32", "TA_BITPAR_PLANES": "2", "TA_BITPAR_THREADS": "64"}, LEV_TESTS),
    ({"TA_NO_SEARCH_FILTER": "1", "TA_SEARCH_KERNEL": "thread"}, SEARCH_TESTS),
    ({"TA_NO_SEARCH_FILTER": "1", "TA_SEARCH_KERNEL": "wave"}, SEARCH_TESTS),
    ({"TA_SEARCH_KERNEL": "thread"}, SEARCH_TESTS),
    ({"TA_NO_SEARCH_FILTER": "1", "TA_SEARCH_KERNEL": "global"}, SEARCH_TESTS),
    ({"TA_SEARCH_FILTER": "myers"}, SEARCH_TESTS),
    ({"TA_SEARCH_FILTER": "pigeon"}, SEARCH_TESTS),
    ({"TA_SEARCH_FILTER": "pigeon", "TA_PIGEON_STAGED": "0"}, SEARCH_TESTS),
    ({"TA_SEARCH_FILTER": "qgram"}, SEARCH_TESTS),
    ({"TA_SEARCH_FILTER": "qgram", "TA_QGRAM_QCAP": "1"}, SEARCH_TESTS),
], ids=["general-band-kernel", "diagonal-extension-kernel-forced", "u16-thread-per-pair-kernel-forced", "u16-thread-per-pair-kernel-on-unit-costs",
        "u16-thread-per-pair-kernel-single-stage", "u16-thread-per-pair-kernel-off", "diagonal-extension-kernel-off", "bitpar-simd-kernel", "bitpar-sliding-table-kernel",
        "bitpar-sliding-table-32bit-on-narrow-bands", "bitpar-block-table-one-pair-per-thread", "bitpar-duo-kernel-without-tile-ordering", "bitpar-duo-kernel-tile-ordered-always", "bitpar-block-table-tile-ordered-always", "bitpar-no-tile-ordering", "length-bucketing-pre-pass", "bitpar-block-table-256-entries",
        "bitpar-block-table-8-blocks", "bitpar-block-table-256-entries-8-blocks",
        "bitpar-table-2plane-kernel", "search-thread-kernel-nofilter", "search-wave-kernel-nofilter",
        "search-thread-kernel-filter", "search-global-rows-kernel-nofilter", "search-myers-filter-forced", "search-pigeonhole-filter-forced",
        "search-pigeonhole-filter-lane-per-segment-loads", "search-qgram-filter-forced-on-small-batches",
        "search-qgram-filter-gives-up-and-falls-back"])
def test_every_kernel_variant_forced(env, select):
    """The dispatchers pick a kernel from the cost model, band width and batch size (bit-parallel vs general banded
    kernel; pre-filter + warp-wavefront vs thread-per-haystack exact search).  These switches force the variants
    the default run does not reach, and the differential tests run again in a subprocess, so that every kernel is
    pinned to the oracle on the same inputs."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.abspath(__file__), "-k",
                        select], env=dict(os.environ, **env), capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
