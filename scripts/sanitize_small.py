"""Small invocation of every kernel added in the second half of round 1, for compute-sanitizer runs:
    compute-sanitizer --tool memcheck  python scripts/sanitize_small.py
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py
Results are checked against the oracle as well (test infrastructure)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as orc  # noqa: E402
import triple_accel_b200 as ta  # noqa: E402
from triple_accel_b200 import synth  # noqa: E402

eng = ta.Engine(0)
# duo (k <= 8), block table C = 16 (k = 16), C = 8 (k = 20 and RDAMERAU k = 16), sliding tables (k = 30, 60); ragged lengths
for length, k, costs in ((128, 8, (1, 1, 0, 0)), (100, 5, (1, 1, 0, 0)), (128, 16, (1, 1, 0, 0)), (77, 20, (1, 1, 0, 0)),
                         (200, 16, (1, 1, 0, 1)), (150, 30, (1, 1, 0, 0)), (150, 60, (1, 1, 0, 1))):
    a, ao, b, bo = synth.mutated_pairs(3000, length, k + 2, seed=length + k, allow_swap=bool(costs[3]))
    got = eng.levenshtein_k_batch(a, ao, b, bo, k, costs)
    assert np.array_equal(got, orc.levenshtein_k_batch(a, ao, b, bo, k, costs, threads=4)), (length, k, costs)
    ra, rao, rb, rbo = synth.random_pairs(3000, length, seed=7)
    got = eng.levenshtein_k_batch(ra, rao, rb, rbo, k, costs)  # early exit path
    assert np.array_equal(got, orc.levenshtein_k_batch(ra, rao, rb, rbo, k, costs, threads=4)), ("R", length, k)
# exponential search (first round k = 16, survivors at 30, 60)
a, ao, b, bo = synth.mutated_pairs(1500, 300, 40, seed=3)
assert np.array_equal(eng.levenshtein_exp_batch(a, ao, b, bo), orc.levenshtein_exp_batch(a, ao, b, bo, threads=4))
# search: exact-piece pre-filter (staged), Myers pre-filter (needle 40), exact wave kernel
for nlen, k in ((32, 3), (16, 2), (40, 4)):
    needle, hay, hoff = synth.needle_haystacks(300, 3000, nlen, plant_frac=0.2, max_edits=k, seed=nlen)
    for st in (0, 1):
        got, goff = eng.levenshtein_search_batch(needle, hay, hoff, k, st)
        want, woff = orc.levenshtein_search_batch(needle, hay, hoff, k, st, threads=4)
        assert np.array_equal(goff, woff) and np.array_equal(got, want), (nlen, k, st)
print("sanitize_small ok: %d launches" % eng.launch_count)
eng.close()
