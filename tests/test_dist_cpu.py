"""world_size-2 gloo tests of the multi-GPU host logic (sharding, result gather, needle broadcast) on the CPU box.
The per-shard compute is the oracle here (no GPU in this container): the test pins the plumbing, not the kernels."""
import os
import socket

import numpy as np
import pytest

import _oracle as orc
from triple_accel_b200 import dist as tdist
from triple_accel_b200 import synth


def test_shard_bounds_balanced_and_contiguous():
    a, ao, b, bo = synth.mutated_pairs(1000, 64, 4, seed=3)
    for world in (1, 2, 3, 8):
        bd = tdist.shard_bounds(ao, bo, world)
        assert bd[0] == 0 and bd[-1] == 1000 and all(x <= y for x, y in zip(bd, bd[1:]))
        sizes = [bd[i + 1] - bd[i] for i in range(world)]
        assert max(sizes) - min(sizes) <= 4
    # ragged: one huge string must not starve the other ranks
    lens = np.array([10] * 50 + [10000] + [10] * 49, np.uint64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    bd = tdist.shard_bounds(off, off, 2)
    assert 45 <= bd[1] <= 52
    # empty batch
    assert tdist.shard_bounds(np.zeros(1, np.uint64), np.zeros(1, np.uint64), 4) == [0, 0, 0, 0, 0]


def test_c_abi_shard_bounds_match_the_contract():
    """ta_shard_bounds = the split ta_init_multi applies inside the library (csrc/multi.cu): contiguous, covers the batch,
    balanced by bytes, tolerant of empty strings, empty batches and more devices than units"""
    import ctypes as C
    from triple_accel_b200 import _ffi
    lib = _ffi.load()

    def bounds(ao, bo, parts):
        out = np.zeros(parts + 1, np.uint64)
        rc = lib.ta_shard_bounds(ao.ctypes.data_as(C.c_void_p), None if bo is None else bo.ctypes.data_as(C.c_void_p),
                                 len(ao) - 1, parts, out.ctypes.data_as(C.c_void_p))
        assert rc == 0
        return [int(x) for x in out]

    a, ao, b, bo = synth.ragged_mutated_pairs(5000, 0, 300, 6, seed=9, templates=5000)
    for parts in (1, 2, 3, 8):
        bd = bounds(ao, bo, parts)
        assert bd[0] == 0 and bd[-1] == 5000 and all(x <= y for x, y in zip(bd, bd[1:]))
        total = int(ao[-1] + bo[-1])
        shares = [int(ao[bd[r + 1]] - ao[bd[r]] + bo[bd[r + 1]] - bo[bd[r]]) for r in range(parts)]
        assert max(shares) - min(shares) <= 2 * 600 + total // 1000, (parts, shares)   # within a pair or two of equal
    lens = np.array([10] * 50 + [10000] + [10] * 49, np.uint64)     # one huge string must not starve the others
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    assert 45 <= bounds(off, off, 2)[1] <= 52
    z = np.zeros(1, np.uint64)
    assert bounds(z, z, 4) == [0, 0, 0, 0, 0]                       # empty batch
    e = np.zeros(8, np.uint64)
    assert bounds(e, e, 3) == [0, 2, 4, 7]                           # all-empty strings: balanced by count
    two = np.array([0, 5, 9], np.uint64)
    bd = bounds(two, None, 8)                                        # more devices than units
    assert bd[0] == 0 and bd[-1] == 2 and all(x <= y for x, y in zip(bd, bd[1:]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, ao, b, bo = synth.mutated_pairs(3001, 96, 10, seed=5)
        full = tdist.sharded_pairs(lambda sa, sao, sb, sbo: orc.levenshtein_k_batch(sa, sao, sb, sbo, 6),
                                   a, ao, b, bo)
        want = orc.levenshtein_k_batch(a, ao, b, bo, 6)
        ok1 = bool(np.array_equal(full, want))
        needle, hay, hoff = synth.needle_haystacks(101, 512, 16, plant_frac=0.2, seed=6)
        my_needle = needle if rank == 0 else np.zeros(3, np.uint8)  # only rank 0 knows the real needle
        m, moff, lo, hi = tdist.sharded_search(lambda nd, h, ho: orc.levenshtein_search_batch(nd, h, ho, 2, 1),
                                               my_needle, hay, hoff)
        wm, wmoff = orc.levenshtein_search_batch(needle, hay, hoff, 2, 1)
        ok2 = bool(np.array_equal(m, wm[int(wmoff[lo]):int(wmoff[hi])]) and
                   np.array_equal(moff, wmoff[lo:hi + 1] - wmoff[lo]))
        q.put((rank, ok1, ok2, lo, hi))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_sharded_levenshtein_and_search():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] for r in res), res
    assert res[0][3] == 0 and res[0][4] == res[1][3] and res[1][4] == 101  # contiguous haystack ranges
