"""Small invocation of every kernel added in round 2, for compute-sanitizer runs:
    compute-sanitizer --tool memcheck  python scripts/sanitize_round2.py
    compute-sanitizer --tool racecheck python scripts/sanitize_round2.py
Tight numpy arrays (no slack after the last string) on purpose: the kernels must not read past a string's last aligned
word.  Results are checked against the oracle as well (test infrastructure)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("TA_SEARCH_FILTER", "qgram")  # the dispatcher takes the q-gram scan from 64 MB per call: force it here
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as orc  # noqa: E402
import triple_accel_b200 as ta  # noqa: E402
from triple_accel_b200 import synth  # noqa: E402

eng = ta.Engine(0)
# diagonal-extension kernel (long strings; transpositions; a 2-letter alphabet = many cooperative slides)
for alpha, costs in ((256, (1, 1, 0, 0)), (2, (1, 1, 0, 1)), (4, (1, 1, 0, 0))):
    a, ao, b, bo = synth.edited_pairs(300, 1024, 2300, 12, seed=alpha, allow_swap=True, alphabet=alpha)
    for k in (3, 16):
        got = eng.levenshtein_k_batch(a, ao, b, bo, k, costs)
        assert np.array_equal(got, orc.levenshtein_k_batch(a, ao, b, bo, k, costs, threads=4)), ("fr", alpha, k)
    assert np.array_equal(eng.levenshtein_exp_batch(a, ao, b, bo, costs), orc.levenshtein_exp_batch(a, ao, b, bo, costs, threads=4))
# thread-per-pair u16 kernel (>= 16384 pairs): affine, transpositions, two-stage register count, ragged incl. empty
a, ao, b, bo = synth.edited_pairs(17000, 0, 90, 8, seed=5, allow_swap=True, alphabet=4)
for costs, k in (((2, 1, 3, 0), 16), ((2, 2, 1, 3), 9), ((1, 1, 0, 1), 6), ((5, 4, 3, 0), 30)):
    got = eng.levenshtein_k_batch(a, ao, b, bo, k, costs)
    assert np.array_equal(got, orc.levenshtein_k_batch(a, ao, b, bo, k, costs, threads=4)), ("diag16", costs, k)
# two pairs per thread behind the tile-local length ordering: ragged lengths incl. empty strings, sizes around a tile,
# an equal-length batch (one-class tiles keep their order); the host-buffer call decides from the offsets, the hint forces
for n_pairs, lo, hi in ((5000, 0, 220), (3073, 96, 160), (2500, 128, 128)):
    a, ao, b, bo = synth.edited_pairs(n_pairs, lo, hi, 8, seed=n_pairs, allow_swap=False)
    want = orc.levenshtein_k_batch(a, ao, b, bo, 8, threads=4)
    want16 = orc.levenshtein_k_batch(a, ao, b, bo, 16, (1, 1, 0, 1), threads=4)
    for hint in (None, True):
        eng.set_length_hint(hint)
        assert np.array_equal(eng.levenshtein_k_batch(a, ao, b, bo, 8), want), ("duo tiled", n_pairs, hint)
        assert np.array_equal(eng.levenshtein_k_batch(a, ao, b, bo, 16, (1, 1, 0, 1)), want16), ("blk tiled", n_pairs, hint)
eng.set_length_hint(None)
# search: weighted costs through the pre-filter, long needle on the global-rows kernel
needle, hay, hoff = synth.needle_haystacks(200, 3000, 32, plant_frac=0.2, max_edits=3, seed=9)
for costs in ((2, 1, 3, 0), (2, 2, 1, 3)):
    got, goff = eng.levenshtein_search_batch(needle, hay, hoff, 6, 1, costs)
    want, woff = orc.levenshtein_search_batch(needle, hay, hoff, 6, 1, costs, threads=4)
    assert np.array_equal(goff, woff) and np.array_equal(got, want), ("search weighted", costs)
# aligned-word 4-gram scan + resolve kernel + split wave items: needles 32 and 64, unaligned ragged haystacks (tight array),
# 256-letter alphabet (the scan's own path) and a 4-letter one (queue overflow -> device-side fallback kernel)
for alpha, nlen, k in ((256, 32, 3), (256, 64, 6), (4, 32, 3), (4, 57, 2)):
    rs = np.random.default_rng(alpha + nlen)
    nd = rs.integers(0, alpha, size=nlen, dtype=np.uint8)
    lens = rs.integers(0, 900, size=400)
    lens[:3] = (0, 3, nlen)
    hoff2 = np.zeros(401, dtype=np.uint64)
    hoff2[1:] = np.cumsum(lens)
    hay2 = rs.integers(0, alpha, size=int(hoff2[-1]), dtype=np.uint8)
    for i in range(3, 400, 5):
        if lens[i] > nlen + 4:
            pos = int(hoff2[i]) + int(rs.integers(0, lens[i] - nlen))
            hay2[pos:pos + nlen] = nd
            hay2[pos + int(rs.integers(0, nlen))] ^= 1
    for st in (0, 1):
        got, goff = eng.levenshtein_search_batch(nd, hay2, hoff2, k, st)
        want, woff = orc.levenshtein_search_batch(nd, hay2, hoff2, k, st, threads=4)
        assert np.array_equal(goff, woff) and np.array_equal(got, want), ("search qgram", alpha, nlen, k, st)
rng = np.random.default_rng(3)
long_needle = rng.integers(1, 5, size=500, dtype=np.uint8)
hay = rng.integers(1, 5, size=40 * 1500, dtype=np.uint8)
hay[3000:3500] = long_needle
hoff = synth.fixed_offsets(40, 1500)
got, goff = eng.levenshtein_search_batch(long_needle, hay, hoff, 30, 1)
want, woff = orc.levenshtein_search_batch(long_needle, hay, hoff, 30, 1, threads=4)
assert np.array_equal(goff, woff) and np.array_equal(got, want), "search long needle"
# traceback grouped by distance (k = u32::MAX on long strings)
a, ao, b, bo = synth.edited_pairs(40, 600, 1500, 6, seed=11)
dist, edits, eoff = eng.levenshtein_k_trace_batch(a, ao, b, bo, 0xFFFFFFFF, (1, 1, 0, 0))
assert np.array_equal(dist, orc.levenshtein_k_batch(a, ao, b, bo, 0xFFFFFFFF, threads=4))
print("sanitize_round2 ok: %d launches" % eng.launch_count)
eng.close()
