"""Model check (CPU, oracle only) of the restart margin the exact search kernel uses on flagged sub-segments
(triple_accel_b200/csrc/search.cu: search_wave_kernel, `warm`).

Claim: for an unanchored search, every end position x (1-based column) with cost <= k inside a sub-segment gets the
reference's (start, end, cost) from a DP that is restarted `needle_len + k / gap_cost + 2` bytes before the
sub-segment.  Why: costs are non-negative, so a cell of cost c <= k is decided by predecessor cells of cost <= c; a
path of cost c to needle row j consumes at most j + c / gap_cost haystack bytes (every haystack byte beyond the
needle's costs a gap), so every candidate that wins OR TIES (the reference breaks ties by match length,
src/levenshtein.rs:1755-1779) lies inside the restarted window, and the candidates the restart sees too expensively
lose in both runs.  Small alphabets make ties the rule.
"""
import random

import pytest

import _oracle as orc

SUB = 128

COSTS = [
    (1, 1, 0, 0),
    (1, 1, 0, 1),
    (2, 1, 3, 0),
    (1, 2, 0, 1),
    (3, 2, 1, 2),
    (1, 1, 2, 0),
]


@pytest.mark.parametrize("costs", COSTS, ids=[str(c) for c in COSTS])
def test_restart_margin_reproduces_every_match_in_the_segment(costs):
    rng = random.Random(hash(costs) & 0xFFFF)
    checked = 0
    for _ in range(120):
        alpha = rng.choice([2, 2, 3, 4])
        n = rng.randint(1, 24)
        needle = bytes(rng.randrange(alpha) for _ in range(n))
        hlen = rng.randint(SUB + 1, 4 * SUB)
        hay = bytearray(rng.randrange(alpha) for _ in range(hlen))
        for _ in range(rng.randint(0, 3)):  # plant a few near-copies
            p = rng.randrange(hlen)
            hay[p:p + n] = needle[: max(0, min(n, hlen - p))]
        hay = bytes(hay)
        gap = costs[1]
        k = rng.randint(0, max(1, n * gap // 2))
        full = orc.levenshtein_search_naive_with_opts(needle, hay, k, 0, costs, False)  # All: every end with cost <= k
        for seg in range((hlen + SUB - 1) // SUB):
            emit_from = seg * SUB
            seg_end = min(emit_from + SUB, hlen)
            warm = n + k // gap + 2
            col0 = max(0, emit_from - warm)
            part = orc.levenshtein_search_naive_with_opts(needle, hay[col0:seg_end], k, 0, costs, False)
            got = sorted((s + col0, e + col0, c) for (s, e, c) in part if e + col0 > emit_from)
            want = sorted((s, e, c) for (s, e, c) in full if emit_from < e <= seg_end)
            assert got == want, (needle, hay, k, costs, seg)
            checked += len(want)
    assert checked > 200
