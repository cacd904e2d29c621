"""Model check of the exact-piece search pre-filter (triple_accel_b200/csrc/lev_bitpar.cu: search_pigeon_kernel), on the
CPU: the two facts the kernel relies on, checked against the scalar oracle on small alphabets where near-matches and
ties are everywhere.

 (i)  coverage: for every end position x with cost <= k there is a needle piece (k + 1 pieces; 2k + 1 with
      transpositions) that occurs EXACTLY in the haystack, ending at byte q, with x - 1 in [q + r - k, q + r + k]
      (r = needle bytes after the piece);
 (ii) verification: the semi-global DP restarted at st = q - fin - k (fin = last needle index of the piece) already
      sees cost <= k at that end, i.e. the candidate's bounded window is enough to confirm it.
"""
import random

import pytest

import _oracle as orc


def _pieces(n, count):
    base, extra = divmod(n, count)
    out, s = [], 0
    for i in range(count):
        ln = base + (1 if i < extra else 0)
        out.append((s, ln))
        s += ln
    return out


@pytest.mark.parametrize("trans", [False, True], ids=["levenshtein", "rdamerau"])
def test_every_match_end_has_a_confirming_piece_occurrence(trans):
    rng = random.Random(5 + trans)
    costs = (1, 1, 0, 1 if trans else 0)
    checked = 0
    for _ in range(250):
        alpha = rng.choice([2, 3, 4, 8])
        n = rng.randrange(4, 33)
        k = rng.randrange(0, 4)
        count = 2 * k + 1 if trans else k + 1
        if count > n:
            continue
        needle = bytes(1 + rng.randrange(alpha) for _ in range(n))
        hay = bytearray(1 + rng.randrange(alpha) for _ in range(rng.randrange(n, 200)))
        for _ in range(rng.randrange(0, 3)):  # plant near-matches
            p = rng.randrange(len(hay) - n + 1)
            s = bytearray(needle)
            for _ in range(rng.randrange(0, k + 1)):
                kind = rng.randrange(4 if trans else 3)
                if kind == 0:
                    s[rng.randrange(len(s))] = 1 + rng.randrange(alpha)
                elif kind == 1:
                    s.insert(rng.randrange(len(s) + 1), 1 + rng.randrange(alpha))
                elif kind == 2 and len(s) > 1:
                    del s[rng.randrange(len(s))]
                elif kind == 3 and len(s) > 1:
                    i = rng.randrange(len(s) - 1)
                    s[i], s[i + 1] = s[i + 1], s[i]
            hay[p:p + len(s)] = s[:max(0, len(hay) - p)]
        hay = bytes(hay)
        ends = {e: c for _, e, c in orc.levenshtein_search_naive_with_opts(needle, hay, k, 0, costs) if e > 0}
        pieces = _pieces(n, count)
        occ = []  # (q = index of the piece's last byte, fin = its last needle index)
        for s0, ln in pieces:
            pat = needle[s0:s0 + ln]
            pos = hay.find(pat)
            while pos >= 0:
                occ.append((pos + ln - 1, s0 + ln - 1))
                pos = hay.find(pat, pos + 1)
        for e in ends:
            xb = e - 1  # byte index of the match end
            ok = False
            for q, fin in occ:
                r = n - 1 - fin
                if not (q + r - k <= xb <= q + r + k):
                    continue
                st = max(0, q - fin - k)
                sub = orc.levenshtein_search_naive_with_opts(needle, hay[st:xb + 1], k, 0, costs)
                if any(ee == xb + 1 - st for _, ee, _ in sub):
                    ok = True
                    break
            assert ok, (needle, hay, k, e)
            checked += 1
    assert checked > 300


@pytest.mark.parametrize("trans", [False, True], ids=["levenshtein", "rdamerau"])
def test_aligned_word_sampling_finds_every_match_end(trans):
    """search_qgram_kernel (pieces of >= 7 bytes): looking only at the 4-byte ALIGNED words of memory and keeping those
    that equal a 4-byte substring of a piece, each such (piece, offset) candidate verified in the same bounded window,
    still confirms every match end -- for every alignment of the haystack in memory."""
    rng = random.Random(15 + trans)
    costs = (1, 1, 0, 1 if trans else 0)
    checked = 0
    for _ in range(200):
        alpha = rng.choice([2, 3, 4, 8])
        k = rng.randrange(0, 3)
        count = 2 * k + 1 if trans else k + 1
        n = rng.randrange(7 * count, min(33, 7 * count + 12)) if 7 * count <= 32 else 0
        if n == 0:
            continue
        needle = bytes(1 + rng.randrange(alpha) for _ in range(n))
        hay = bytearray(1 + rng.randrange(alpha) for _ in range(rng.randrange(n, 160)))
        for _ in range(rng.randrange(1, 3)):  # plant near-matches
            p = rng.randrange(len(hay) - n + 1)
            s = bytearray(needle)
            for _ in range(rng.randrange(0, k + 1)):
                kind = rng.randrange(4 if trans else 3)
                if kind == 0:
                    s[rng.randrange(len(s))] = 1 + rng.randrange(alpha)
                elif kind == 1:
                    s.insert(rng.randrange(len(s) + 1), 1 + rng.randrange(alpha))
                elif kind == 2 and len(s) > 1:
                    del s[rng.randrange(len(s))]
                elif kind == 3 and len(s) > 1:
                    i = rng.randrange(len(s) - 1)
                    s[i], s[i + 1] = s[i + 1], s[i]
            hay[p:p + len(s)] = s[:max(0, len(hay) - p)]
        hay = bytes(hay)
        ends = {e for _, e, c in orc.levenshtein_search_naive_with_opts(needle, hay, k, 0, costs) if e > 0}
        grams = []  # (4 bytes, bytes from the word's first byte to the piece's last byte, fin)
        for s0, ln in _pieces(n, count):
            assert ln >= 7
            for o in range(ln - 3):
                grams.append((needle[s0 + o:s0 + o + 4], ln - 1 - o, s0 + ln - 1))
        mis = rng.randrange(4)  # address of hay[0] modulo 4
        found = set()
        for x in range(len(hay) - 3):
            if (mis + x) % 4:
                continue  # not an aligned word
            word = hay[x:x + 4]
            for g, to_end, fin in grams:
                if g != word:
                    continue
                q = x + to_end
                if q >= len(hay):
                    continue
                r = n - 1 - fin
                lo, hi = max(0, q + r - k), min(len(hay) - 1, q + r + k)
                st = max(0, q + 1 - (fin + 1 + k))
                sub = orc.levenshtein_search_naive_with_opts(needle, hay[st:hi + 1], k, 0, costs)
                for _, ee, _ in sub:
                    if ee > 0 and lo <= st + ee - 1 <= hi:
                        found.add(st + ee)
        # the filter flags 128-byte sub-segments; here we ask for more: every single end is confirmed
        assert ends <= found, (needle, hay, k, mis, sorted(ends - found))
        checked += len(ends)
    assert checked > 200
