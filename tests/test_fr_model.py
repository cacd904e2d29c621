"""Executable model of the furthest-reaching ("diagonal extension") unit-cost kernel
(triple_accel_b200/csrc/lev_fr.cu: lev_fr_kernel) in plain Python, checked against the scalar oracle on the CPU.

The kernel answers levenshtein_naive_k_with_opts (reference src/levenshtein.rs:376-545) for LEVENSHTEIN_COSTS and
RDAMERAU_COSTS without filling the band: level e = 0, 1, ... holds, per diagonal c = j - i, the furthest row
FR_e[c] with D[row][row + c] <= e; a level is derived from the previous one (substitution / the two gaps /
restricted transposition) and extended along runs of equal bytes.  The model pins the parts that are easy to get
wrong: the per-level diagonal range (Ukkonen trim against the target diagonal, |c| <= e, c >= -|a|), the clamp to
the matrix, which stale entries of the two ping-pong arrays may be read, and the transposition candidate.  Any
change to the kernel's recurrence should be made here first."""
import random

import _oracle as orc

NEG = -(1 << 30)


def fr_distance(a: bytes, b: bytes, k: int, trans: bool):
    """None or the distance; the contract of levenshtein_naive_k_with_opts for unit costs"""
    if len(a) > len(b):
        a, b = b, a
    m, n = len(a), len(b)
    diff = n - m
    max_k = min(k, n)
    if diff > max_k:
        return None
    if m == 0:
        return n
    slots = 2 * max_k + 3  # slot = c + max_k + 1: one never-written guard slot on each side
    fr = [[NEG] * slots, [NEG] * slots]
    for e in range(max_k + 1):
        cur, prev = fr[e & 1], fr[(e & 1) ^ 1]
        lo = max(-e, -m, diff - (max_k - e))
        hi = min(e, diff + (max_k - e))
        for c in range(lo, hi + 1):
            s = c + max_k + 1
            lim = min(m, n - c)
            if e == 0:
                t = 0
            else:
                mid = prev[s]
                t = max(mid + 1, prev[s - 1], prev[s + 1] + 1)
                if trans and mid >= 0 and mid + 2 <= lim and a[mid] == b[mid + c + 1] and a[mid + 1] == b[mid + c]:
                    t = max(t, mid + 2)
            assert t >= max(0, -c), (t, c, e)
            t = min(t, lim)
            while t < lim and a[t] == b[t + c]:
                t += 1
            cur[s] = t
        if cur[diff + max_k + 1] >= m:
            return e
    return None


def rand_pair(rng, alpha, max_len):
    la = rng.randrange(0, max_len + 1)
    a = bytes(rng.randrange(alpha) for _ in range(la))
    if rng.random() < 0.7:  # b = a after a few edits (swaps included)
        s = bytearray(a)
        for _ in range(rng.randrange(0, 7)):
            kind = rng.randrange(4)
            if kind == 0 and s:
                s[rng.randrange(len(s))] = rng.randrange(alpha)
            elif kind == 1:
                s.insert(rng.randrange(len(s) + 1), rng.randrange(alpha))
            elif kind == 2 and s:
                del s[rng.randrange(len(s))]
            elif kind == 3 and len(s) > 1:
                p = rng.randrange(len(s) - 1)
                s[p], s[p + 1] = s[p + 1], s[p]
        b = bytes(s)
    else:
        b = bytes(rng.randrange(alpha) for _ in range(rng.randrange(0, max_len + 1)))
    return a, b


def check(trans, cases, seed):
    rng = random.Random(seed)
    costs = (1, 1, 0, 1) if trans else (1, 1, 0, 0)
    for _ in range(cases):
        a, b = rand_pair(rng, rng.choice((2, 3, 4, 26)), rng.choice((6, 12, 40)))
        k = rng.choice((0, 1, 2, 3, 5, 8, 16, 100, 0xFFFFFFFF))
        want = orc.levenshtein_naive_k_with_opts(a, b, k, False, costs)
        want = None if want is None else want[0]
        assert fr_distance(a, b, k, trans) == want, (a, b, k, trans)


def test_fr_model_levenshtein():
    check(False, 6000, 11)


def test_fr_model_rdamerau():
    check(True, 6000, 12)


def test_fr_model_regressions():
    # the scalar-vs-SIMD regression cases of SURVEY Appendix B (scalar answers)
    assert fr_distance(b"yxy", b"yx", 10, True) == 1
    assert fr_distance(b"x\0", b"x", 10, True) == 1
    assert fr_distance(b"zzzyxy\0", b"yx\0", 10, True) == 4
    assert fr_distance(b"xyyzzyzy", b"yyxyzzyz", 10, True) == 3
