"""Executable model (CPU, no GPU code involved) of the tile-local length ordering of lev_bitpar_duo_tiled_kernel
(triple_accel_b200/csrc/lev_bitpar.cu): the invariants the kernel's pairing relies on.

 * classes are counted, every class starts at an EVEN slot, an odd class leaves its last slot empty (0xFFFF);
 * round r gives thread t the slots 2 (128 r + t) and + 1: the two are of one class or the second is empty, so a
   thread never straddles two classes (that sent its whole warp down the single-pair path);
 * every pair of the tile appears exactly once; the slots in use are at most the pairs + one per class;
 * the CTA split: contiguous rounds of 256 pairs, every CTA within one round of the others, tiles of <= 3072 pairs.
"""
import random

DUO_TILE, CLASSES, EMPTY = 3072, 64, 0xFFFF


def order_tile(lengths):
    cls = [min(max(la, lb) >> 4, CLASSES - 1) for la, lb in lengths]
    cnt = [0] * CLASSES
    for c in cls:
        cnt[c] += 1
    if max(cnt) == len(lengths):
        return list(range(len(lengths))), len(lengths)  # one class: identity
    start, pos = [0] * CLASSES, 0
    for c in range(CLASSES):
        start[c] = pos
        pos += (cnt[c] + 1) & ~1
    order = [None] * pos
    for c in range(CLASSES):
        if cnt[c] & 1:
            order[start[c] + cnt[c]] = EMPTY
    cur = start[:]
    for i, c in enumerate(cls):  # the kernel places by atomicAdd: any order inside a class
        order[cur[c]] = i
        cur[c] += 1
    return order, pos


def test_every_thread_gets_one_class_and_every_pair_once():
    rng = random.Random(5)
    for trial in range(300):
        n = rng.choice([1, 2, 3, 255, 256, 257, 1000, DUO_TILE - 1, DUO_TILE])
        lo, hi = rng.choice([(0, 40), (96, 160), (128, 128), (0, 1200), (15, 17)])
        lengths = [(rng.randint(lo, hi), rng.randint(lo, hi)) for _ in range(n)]
        order, plen = order_tile(lengths)
        assert plen <= n + CLASSES and len(order) == plen and None not in order
        assert sorted(x for x in order if x != EMPTY) == list(range(n))
        for s in range(0, plen, 2):  # slots 2s, 2s + 1 of one thread
            a = order[s]
            b = order[s + 1] if s + 1 < plen else EMPTY
            assert a != EMPTY
            if b != EMPTY:
                ca = min(max(lengths[a]) >> 4, CLASSES - 1)
                cb = min(max(lengths[b]) >> 4, CLASSES - 1)
                assert ca == cb, (trial, s)


def test_cta_split_covers_the_batch_in_balanced_contiguous_rounds():
    rng = random.Random(6)
    for trial in range(300):
        n = rng.choice([1, 255, 256, 257, 100_000, 1_000_000, rng.randint(1, 3_000_000)])
        ctas = rng.choice([1, 3, 444])
        rounds = (n + 255) // 256
        grid = min(rounds, ctas)
        rq, rrem = divmod(rounds, grid)
        covered, sizes = 0, []
        for b in range(grid):
            r_lo = b * rq + min(b, rrem)
            r_cnt = rq + (1 if b < rrem else 0)
            assert r_lo * 256 == covered or r_lo * 256 >= n
            n_tiles = (r_cnt + DUO_TILE // 256 - 1) // (DUO_TILE // 256)
            for t in range(n_tiles):
                t_lo, t_hi = r_lo + r_cnt * t // n_tiles, r_lo + r_cnt * (t + 1) // n_tiles
                base, end = t_lo * 256, min(t_hi * 256, n)
                assert 0 < end - base <= DUO_TILE and base == covered
                covered = end
            sizes.append(r_cnt)
        assert covered == n and max(sizes) - min(sizes) <= 1
