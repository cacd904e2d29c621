"""Model check of the Rust shim's lazy `SearchType::All` iterator (bindings/rust/triple_accel/src/levenshtein.rs: LazyAll;
the reference's iterator is lazy, src/levenshtein.rs:2282, 2448): searching a haystack chunk by chunk, each chunk
restarted 2 |needle| + start_gap / gap + 2 bytes early and keeping only the matches that end inside the chunk, gives
exactly the match list of one search over the whole haystack.  Checked against the oracle with tiny chunks on small
alphabets (where length ties are the rule), unit / weighted / affine / transposition costs."""
import random

import _oracle as orc


def lazy_all(needle, hay, k, costs, chunk):
    out, lo = [], 0
    warm = 2 * len(needle) + costs[2] // costs[1] + 2
    while lo <= len(hay):
        hi = min(len(hay), lo + chunk)
        frm = max(0, lo - warm)
        part = orc.levenshtein_search_naive_with_opts(needle, hay[frm:hi], k, 0, costs, False)
        for (s, e, c) in part:
            if lo == 0 or e + frm > lo:
                out.append((s + frm, e + frm, c))
        lo = hi + 1 if hi == len(hay) else hi
    return out


def test_lazy_all_equals_one_search():
    rng = random.Random(4)
    models = [(1, 1, 0, 0), (1, 1, 0, 1), (2, 1, 3, 0), (2, 2, 1, 3), (1, 2, 5, 0), (3, 1, 0, 0)]
    for it in range(1500):
        alpha = rng.choice([2, 3, 4])
        needle = bytes(rng.randrange(1, alpha + 1) for _ in range(rng.randrange(1, 9)))
        hay = bytes(rng.randrange(1, alpha + 1) for _ in range(rng.randrange(0, 200)))
        costs = rng.choice(models)
        k = rng.randrange(0, 8)
        want = orc.levenshtein_search_naive_with_opts(needle, hay, k, 0, costs, False)
        for chunk in (16, 37):
            assert lazy_all(needle, hay, k, costs, chunk) == want, (needle, hay, k, costs, chunk)
