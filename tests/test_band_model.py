"""Executable model of the lane algorithm of the general banded anti-diagonal kernel (triple_accel_b200/csrc/lev_band.cu:
lev_band_kernel) in numpy, checked against the scalar oracle on the CPU: the host-side band pre-pass (Ukkonen band for
weighted / affine costs, + one margin diagonal each side for transpositions), the parity-alternating neighbour
exchange (left neighbour on even, upper neighbour on odd anti-diagonals: one shuffle per step), sender-side
pre-minimised gap candidates min(D + open, H + gap), the boundary overrides and the final cell.  Any change to the
kernel's recurrence should be made here first."""
import random

import numpy as np

import _oracle as orc

INF = 0x3FFFFFFF

def band(m, n, k, costs):
    """host-side pre-pass: a shorter (m), b longer (n). returns None (=NONE) or (max_k, dlo, W)"""
    mism, gap, sgap, tc = costs
    diff = n - m
    max_k = min(m * mism, 2 * m * gap + (0 if m == 0 else sgap + (sgap if n == m else 0)))
    max_k = min(k, max_k + diff * gap + (0 if n == m else sgap))
    unit_k = max(max_k - sgap, 0) // gap
    if diff > unit_k:
        return None
    num = max_k - 2 * sgap - diff * gap
    e = num // (2 * gap) if num >= 0 else 0
    if tc:  # transposition reads the match flags of the two neighbouring diagonals
        return max_k, -e - 1, diff + 2 * e + 3
    return max_k, -e, diff + 2 * e + 1

def k3(a, b, k, costs, G=32, C=1):
    if len(a) > len(b): a, b = b, a
    m, n = len(a), len(b)
    mism, gap, sgap, tc = costs
    r = band(m, n, k, costs)
    if r is None: return None
    max_k, dlo, W = r
    if m == 0:
        d = n * gap + (sgap if n else 0)
        return d if d <= max_k else None
    assert W <= 2 * G * C, (W, G, C)
    og = sgap + gap
    L = G * C   # cells per step
    cell = np.arange(L)            # cell index = t*C + c
    D1 = np.full(L, INF, np.int64); D2 = D1.copy(); D3 = D1.copy(); D4 = D1.copy()
    oH1 = D1.copy(); oV1 = D1.copy(); m1 = np.zeros(L, bool)
    res = None
    def ch(s, idx):
        idx = np.clip(idx, 0, len(s) - 1)
        return np.frombuffer(s, np.uint8)[idx]
    for s in range(0, m + n + 1):
        p = (s - dlo) & 1
        d = dlo + 2 * cell + p
        # (s-d) is even by construction
        i = (s - d) // 2; j = (s + d) // 2
        if p == 0:
            h = np.concatenate(([INF], oH1[:-1])); mL = np.concatenate(([False], m1[:-1]))
            v = oV1; mU = m1
        else:
            h = oH1; mL = m1
            v = np.concatenate((oV1[1:], [INF])); mU = np.concatenate((m1[1:], [False]))
        eq = ch(a, i - 1) == ch(b, j - 1)
        sub = D2 + np.where(eq, 0, mism)
        D = np.minimum(np.minimum(sub, h), v)
        if tc:
            t = D4 + tc
            D = np.where(mL & mU, np.minimum(D, t), D)
        H = h; V = v
        # boundary override
        bi = (i == 0) & (j >= 0); bj = (j == 0) & (i >= 0)
        bval = np.where(bi, j * gap + np.where(j > 0, sgap, 0), i * gap + np.where(i > 0, sgap, 0))
        isb = bi | bj
        D = np.where(isb, bval, D)
        H = np.where(isb, INF, H); V = np.where(isb, INF, V)
        oH = np.minimum(D + og, H + gap); oV = np.minimum(D + og, V + gap)
        D = np.minimum(D, INF); oH = np.minimum(oH, INF); oV = np.minimum(oV, INF)   # model only: keep bounded
        D4, D3, D2, D1 = D3, D2, D1, D
        oH1, oV1, m1 = oH, oV, eq
        if s == m + n:
            idx = (n - m - dlo - p) // 2
            res = int(D[idx])
    return res if res <= max_k else None



def test_lane_model_matches_the_oracle():
    rng = random.Random(7)
    models = [(1, 1, 0, 0), (1, 1, 0, 1), (1, 1, 2, 0), (2, 1, 2, 0), (2, 3, 0, 0), (3, 1, 0, 0), (2, 2, 1, 3),
              (3, 2, 0, 2), (1, 1, 1, 1)]
    checked = 0
    for it in range(2500):
        alpha = rng.choice([2, 3, 4, 26])
        la, lb = rng.randrange(0, 20), rng.randrange(0, 20)
        a = bytes(rng.randrange(alpha) for _ in range(la))
        b = bytes(rng.randrange(alpha) for _ in range(lb))
        costs = models[it % len(models)]
        k = rng.choice([0, 1, 2, 3, 5, 8, 16, 100, 0xFFFFFFFF])
        G, C = rng.choice([(32, 1), (8, 4), (4, 8), (16, 2)])
        want = orc.levenshtein_naive_k_with_opts(a, b, k, False, costs)
        want = None if want is None else want[0]
        try:
            got = k3(a, b, k, costs, G, C)
        except AssertionError:
            continue  # band wider than this (G, C) lane geometry holds: the dispatcher picks a wider one
        assert got == want, (a, b, k, costs, G, C, want, got)
        checked += 1
    assert checked > 2000
