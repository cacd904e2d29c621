"""Randomised differential run of the search path against the scalar oracle (test infrastructure): needle lengths
1..64, k, alphabets from 2 to 256 symbols, unit / weighted / affine / transposition costs, ragged haystacks (empty,
shorter than the needle, several KB), All and Best, anchored now and then.  Covers every pre-filter (q-gram scan with
queue / resolve / 16-byte granules, its device-side fallback, shift-and, Myers) and both exact kernels as the
dispatcher picks them.
usage: python scripts/stress_search.py [seconds, default 120] [seed]"""
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as orc  # noqa: E402
import triple_accel_b200 as ta  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
eng = ta.Engine(0)
COSTS = [(1, 1, 0, 0), (1, 1, 0, 0), (1, 1, 0, 0), (1, 1, 0, 1), (2, 1, 3, 0), (2, 2, 0, 0), (1, 2, 0, 1), (3, 2, 1, 2), (1, 1, 2, 0)]
t_end = time.time() + budget
cases = matches = 0
while time.time() < t_end:
    alpha = rng.choice([2, 4, 4, 20, 256, 256, 256])
    nlen = rng.choice([1, 2, 7, 13, 14, 21, 28, 31, 32, 33, 40, 48, 57, 63, 64, rng.randint(1, 64)])
    costs = rng.choice(COSTS)
    needle = bytes(rng.randrange(alpha) for _ in range(nlen))
    if rng.random() < 0.15:  # a needle that repeats itself: one word equals several 4-grams
        unit = needle[: rng.choice([1, 2, 4, 5])]
        needle = (unit * (nlen // len(unit) + 1))[:nlen]
    n_hay = rng.choice([1, 3, 40, 300, 1500])
    max_h = rng.choice([40, 300, 700, 5000]) if n_hay < 1000 else rng.choice([40, 300, 700])
    hays = []
    for _ in range(n_hay):
        h = bytearray(rng.randrange(alpha) for _ in range(rng.randrange(0, max_h)))
        for _ in range(rng.randrange(0, 3)):
            if len(h) > nlen + 2:
                pos = rng.randrange(len(h) - nlen)
                m = bytearray(needle)
                for _ in range(rng.randrange(0, 5)):
                    kind = rng.randrange(4)
                    if kind == 0 and m:
                        m[rng.randrange(len(m))] = rng.randrange(alpha)
                    elif kind == 1:
                        m.insert(rng.randrange(len(m) + 1), rng.randrange(alpha))
                    elif kind == 2 and m:
                        del m[rng.randrange(len(m))]
                    elif kind == 3 and len(m) > 1:
                        q = rng.randrange(len(m) - 1)
                        m[q], m[q + 1] = m[q + 1], m[q]
                h[pos:pos + nlen] = m[:nlen]
        hays.append(bytes(h))
    skew = rng.randrange(0, 16)  # haystack bytes start at any alignment of the host buffer
    hay = np.frombuffer(bytes(skew) + b"".join(hays), dtype=np.uint8)
    hoff = np.zeros(n_hay + 1, dtype=np.uint64)
    hoff[0] = skew
    hoff[1:] = skew + np.cumsum([len(h) for h in hays])
    gap = costs[1]
    for k in {0, rng.randint(0, max(1, nlen // 7)), rng.randint(0, nlen * gap + 2), nlen // 4}:
        st = rng.randrange(2)
        anchored = rng.random() < 0.1
        try:
            want, woff = orc.levenshtein_search_batch(needle, hay, hoff, k, st, costs, anchored, threads=8)
        except AssertionError:  # the reference panics on these arguments: so must the drop-in
            try:
                eng.levenshtein_search_batch(needle, hay, hoff, k, st, costs, anchored)
            except AssertionError:
                continue
            print("MISSING PANIC", dict(nlen=nlen, costs=costs, k=k))
            sys.exit(1)
        got, goff = eng.levenshtein_search_batch(needle, hay, hoff, k, st, costs, anchored)
        if not (np.array_equal(goff, woff) and np.array_equal(got, want)):
            print("MISMATCH", dict(alpha=alpha, nlen=nlen, costs=costs, k=k, st=st, anchored=anchored, n_hay=n_hay, max_h=max_h,
                                   needle=needle))
            sys.exit(1)
        cases += 1
        matches += len(want)
print("stress_search ok: %d calls, %d matches checked, %d kernel launches" % (cases, matches, eng.launch_count))
eng.close()
