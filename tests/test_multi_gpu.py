"""Multi-GPU parity (needs >= 2 visible CUDA devices; skipped on a 1-GPU box) and host-thread concurrency.

 * ta_init_multi: ONE context over several devices; every host-buffer batch call is cut into byte-balanced ranges
   inside the library, results must equal the oracle's exactly as on one device; the needle of a
   search must travel by ncclBroadcast.
 * one process per GPU under torch.distributed/NCCL (how bench.py scales): shards through triple_accel_b200.dist,
   real CUDA compute per rank, gathered output equals the oracle's; the needle is known to rank 0 only.
 * concurrency contract of include/triple_accel_b200.h: distinct contexts run concurrently (also on one device), one
   context may be shared by threads.
"""
import os
import socket
import subprocess
import sys
import threading

import numpy as np
import pytest

import _oracle as orc
from triple_accel_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def meng():
    if _ndev() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    import triple_accel_b200 as ta
    os.environ["TA_MULTI_MIN_BYTES"] = "2048"  # small test batches must still be split over every device
    e = ta.Engine(devices=list(range(min(_ndev(), 8))))
    yield e
    e.close()
    os.environ.pop("TA_MULTI_MIN_BYTES", None)


def test_multi_ctx_pairs(meng):
    a, ao, b, bo = synth.mutated_pairs(30011, 128, 10, seed=21, allow_swap=True)
    for costs in ((1, 1, 0, 0), (1, 1, 0, 1), (2, 1, 3, 0)):
        got = meng.levenshtein_k_batch(a, ao, b, bo, 8, costs)
        assert np.array_equal(got, orc.levenshtein_k_batch(a, ao, b, bo, 8, costs, threads=8))
    ha, hao, hb, hbo = synth.hamming_pairs(20000, 64, seed=22)
    assert np.array_equal(meng.hamming_batch(ha, hao, hb, hbo), orc.hamming_batch(ha, hao, hb, hbo))
    a, ao, b, bo = synth.mutated_pairs(3000, 1024, 4, seed=23, exact_edits=True)
    assert np.array_equal(meng.levenshtein_exp_batch(a, ao, b, bo), orc.levenshtein_exp_batch(a, ao, b, bo, threads=8))
    assert meng.launch_count > 0


def test_multi_ctx_ragged_and_tiny(meng):
    # ragged lengths incl. empty strings; a batch smaller than the device count; a batch of one
    a, ao, b, bo = synth.ragged_mutated_pairs(5003, 0, 300, 6, seed=24, templates=5003)
    got = meng.levenshtein_k_batch(a, ao, b, bo, 6)
    assert np.array_equal(got, orc.levenshtein_k_batch(a, ao, b, bo, 6, threads=8))
    try:  # the length hint reaches every device of the context and never changes a result
        for hint in (True, False):
            meng.set_length_hint(hint)
            assert np.array_equal(meng.levenshtein_k_batch(a, ao, b, bo, 6), got)
    finally:
        meng.set_length_hint(None)
    for n in (1, 2, 3):
        sa, sao, sb, sbo = synth.mutated_pairs(n, 40, 3, seed=n)
        assert np.array_equal(meng.levenshtein_k_batch(sa, sao, sb, sbo, 5), orc.levenshtein_k_batch(sa, sao, sb, sbo, 5))
    assert meng.levenshtein(b"kitten", b"sitting") == 3


def test_multi_ctx_search_broadcasts_needle_with_nccl(meng):
    needle, hay, hoff = synth.needle_haystacks(4000, 2048, 32, plant_frac=0.05, seed=25)
    before = meng.needle_broadcasts
    for st in (0, 1):
        got, goff = meng.levenshtein_search_batch(needle, hay, hoff, 3, st)
        want, woff = orc.levenshtein_search_batch(needle, hay, hoff, 3, st, threads=8)
        assert np.array_equal(goff, woff) and np.array_equal(got, want)
    assert meng.uses_nccl, "libnccl.so.2 could not be loaded / ncclCommInitAll failed: needle copied per device"
    assert meng.needle_broadcasts == before + 2
    got, goff = meng.hamming_search_batch(needle, hay, hoff, 3, 0)
    want, woff = orc.hamming_search_batch(needle, hay, hoff, 3, 0, threads=8)
    assert np.array_equal(goff, woff) and np.array_equal(got, want)


def test_multi_ctx_traceback_and_errors(meng):
    a, ao, b, bo = synth.mutated_pairs(2000, 96, 6, seed=26, allow_swap=True)
    dist, edits, eoff = meng.levenshtein_k_trace_batch(a, ao, b, bo, 8, (1, 1, 0, 1))
    assert np.array_equal(dist, orc.levenshtein_k_batch(a, ao, b, bo, 8, (1, 1, 0, 1), threads=8))
    for i in (0, 1, 999, 1000, 1001, 1999):
        sa, sb = bytes(a[int(ao[i]):int(ao[i + 1])]), bytes(b[int(bo[i]):int(bo[i + 1])])
        want = orc.levenshtein_naive_k_with_opts(sa, sb, 8, True, (1, 1, 0, 1))
        got = [tuple(int(x) for x in e) for e in edits[int(eoff[i]):int(eoff[i + 1])]]
        assert (want is None and got == []) or (want is not None and got == [tuple(e) for e in want[1]])
    # a length mismatch in the LAST shard is still the reference's panic
    ha, hao, hb, hbo = synth.hamming_pairs(9000, 64, seed=27)
    hbo = hbo.copy()
    hbo[-1] -= 1
    with pytest.raises(AssertionError):
        meng.hamming_batch(ha, hao, hb[:-1], hbo)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


RANK_SCRIPT = r"""
import os, sys
import numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import torch, torch.distributed as dist
import _oracle as orc
import triple_accel_b200 as ta
from triple_accel_b200 import dist as tdist, synth
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = ta.Engine(local)
a, ao, b, bo = synth.mutated_pairs(40001, 128, 10, seed=31, allow_swap=True)
full = tdist.sharded_pairs(lambda sa, sao, sb, sbo: eng.levenshtein_k_batch(sa, sao, sb, sbo, 8, (1, 1, 0, 1)), a, ao, b, bo)
assert np.array_equal(full, orc.levenshtein_k_batch(a, ao, b, bo, 8, (1, 1, 0, 1), threads=4)), "sharded distances"
needle, hay, hoff = synth.needle_haystacks(2001, 2048, 32, plant_frac=0.05, seed=32)
mine = needle if rank == 0 else np.zeros(5, np.uint8)   # only rank 0 knows the needle: NCCL broadcast
m, moff, lo, hi = tdist.sharded_search(lambda nd, h, ho: eng.levenshtein_search_batch(nd, h, ho, 3, 1), mine, hay, hoff)
wm, wmoff = orc.levenshtein_search_batch(needle, hay, hoff, 3, 1, threads=4)
assert np.array_equal(m, wm[int(wmoff[lo]):int(wmoff[hi])]) and np.array_equal(moff, wmoff[lo:hi + 1] - wmoff[lo]), "sharded search"
assert eng.launch_count > 0
print("rank %%d ok: pairs %%d..%%d" %% (rank, lo, hi))
dist.destroy_process_group()
"""


def test_world2_nccl_one_process_per_gpu(tmp_path):
    """the torchrun layout bench.py scales with: 2 ranks, 2 GPUs, NCCL; gathered output == oracle"""
    if _ndev() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    script = tmp_path / "rank.py"
    script.write_text(RANK_SCRIPT % {"root": ROOT})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.count(" ok: pairs") == 2, r.stdout[-2000:] + r.stderr[-3000:]


def test_threads_two_contexts_and_one_shared_context():
    """include/triple_accel_b200.h: distinct contexts run concurrently (here: on one device), and any number of
    threads may share one context (calls are serialised inside): every result equals the oracle's."""
    import triple_accel_b200 as ta
    e1, e2 = ta.Engine(0), ta.Engine(0)
    sets = []
    for seed in range(4):
        a, ao, b, bo = synth.mutated_pairs(20000 + 37 * seed, 128, 10, seed=40 + seed, allow_swap=True)
        needle, hay, hoff = synth.needle_haystacks(500, 1024, 24, plant_frac=0.05, seed=50 + seed)
        sets.append((a, ao, b, bo, orc.levenshtein_k_batch(a, ao, b, bo, 8, (1, 1, 0, 1), threads=4),
                     needle, hay, hoff, orc.levenshtein_search_batch(needle, hay, hoff, 3, 1, threads=4)))
    errors = []

    def work(eng, which, reps):
        try:
            for r in range(reps):
                a, ao, b, bo, want, needle, hay, hoff, (wm, wmo) = sets[(which + r) % 4]
                got = eng.levenshtein_k_batch(a, ao, b, bo, 8, (1, 1, 0, 1))
                if not np.array_equal(got, want):
                    errors.append(("dist", which, r))
                m, mo = eng.levenshtein_search_batch(needle, hay, hoff, 3, 1)
                if not (np.array_equal(m, wm) and np.array_equal(mo, wmo)):
                    errors.append(("search", which, r))
        except Exception as ex:  # noqa: BLE001
            errors.append(("exception", which, repr(ex)))

    # two threads, two contexts; plus two more threads sharing the first context
    ts = [threading.Thread(target=work, args=(e1, 0, 6)), threading.Thread(target=work, args=(e2, 1, 6)),
          threading.Thread(target=work, args=(e1, 2, 6)), threading.Thread(target=work, args=(e1, 3, 6))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    e1.close()
    e2.close()
    assert not errors, errors[:5]
