#!/usr/bin/env python3
"""bench.py -- throughput of the hot path on B200 (see DESIGN.md "Measurement").

A "step" is one pass of the k-banded Levenshtein kernel over one batch of synthetic pairs.
Default workload = BASELINE.json configs[1]: levenshtein_simd_k, k = 8, 1 M pairs, len 128, unit costs, on the
"matching" set M (b = a after U[0,8] random edits, so no pair can be rejected early).

  python bench.py --gpus 1 --steps 20 --warmup 5            # our CUDA path (one JSON line)
  python bench.py --impl reference --steps 3 --warmup 1     # the reference's CPU algorithm (oracle port) on host cores
  torchrun --nproc-per-node N bench.py --gpus N ...          # one rank per GPU, pairs sharded, no data-path collective

value  = GCUPS (nominal band cells of the reference's scalar banded DP, (2k+1)n - k^2 per pair) with inputs
         resident in HBM, timed with CUDA events on the launching stream, max over ranks.
e2e    = the same metric through the host-buffer C-ABI call (pinned host inputs, H2D + kernel + D2H in the timed
         region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (op, n_pairs, len, k, costs, description)
    "lev_k8_len128": ("lev_k", 1_000_000, 128, 8, (1, 1, 0, 0),
                      "levenshtein_simd_k k=8, 1M pairs len=128, unit costs, set M (b = a after U[0,8] edits)"),
    "lev_k16_len128": ("lev_k", 1_000_000, 128, 16, (1, 1, 0, 0),
                       "levenshtein_simd_k k=16, 1M pairs len=128, unit costs, set M (b = a after U[0,16] edits)"),
    "lev_k8_len128_R": ("lev_k_R", 1_000_000, 128, 8, (1, 1, 0, 0),
                        "levenshtein_simd_k k=8, 1M pairs len=128, unit costs, set R (b independent of a: every pair is "
                        "None; the kernel leaves a pair once its final-diagonal value exceeds k)"),
    "lev_k16_len128_R": ("lev_k_R", 1_000_000, 128, 16, (1, 1, 0, 0),
                         "levenshtein_simd_k k=16, 1M pairs len=128, unit costs, set R (b independent of a)"),
    "lev_k8_ragged96_160": ("lev_k_ragged", 1_000_000, 128, 8, (1, 1, 0, 0),
                            "levenshtein_simd_k k=8, 1M pairs, |a| ~ U[96,160] (mean 128), unit costs, set M: neighbouring "
                            "pairs have unrelated lengths (arbitrary alignment, ragged tails, few pairs can share a thread)"),
    "rdamerau_k16_len512": ("lev_k", 1_000_000, 512, 16, (1, 1, 0, 1),
                            "RDAMERAU_COSTS k=16, 1M pairs len=512, set M (U[0,16] edits incl. swaps)"),
    "lev_k16_len4096": ("lev_k", 262_144, 4096, 16, (1, 1, 0, 0),
                        "levenshtein_simd_k k=16, 256Ki pairs len=4096, unit costs, set M"),
    "affine_k16_len128": ("lev_k", 1_000_000, 128, 16, (2, 1, 3, 0),
                          "levenshtein_simd_k_with_opts k=16, 1M pairs len=128, EditCosts(2,1,3,None) (general kernel)"),
    "lev_k60_len1024": ("lev_k", 262_144, 1024, 60, (1, 1, 0, 0),
                        "levenshtein_simd_k k=60, 256Ki pairs len=1024, unit costs, set M (64-row window kernel)"),
    "exp_len1024": ("exp", 1_000_000, 1024, 30, (1, 1, 0, 0),
                    "levenshtein_exp, 1M pairs len=1024, exactly 4 random edits (first round k=30 succeeds)"),
    "search_n32_h4096": ("search", 100_000, 4096, 3, (1, 1, 0, 0),
                         "levenshtein_search needle len=32 over 100k haystacks len=4096, k=3, Best, 1% planted hits"),
    "hamming_len64": ("hamming", 10_000, 64, 0, (1, 1, 0, 0), "hamming, 10k pairs len=64 (plumbing case)"),
    "hamming_len4096": ("hamming", 262_144, 4096, 0, (1, 1, 0, 0), "hamming, 256Ki pairs len=4096"),
}


def nominal_cells(length, k):
    """cells the reference's scalar banded DP visits for |a| = |b| = length, half-width k"""
    u = min(k, length)
    return (2 * u + 1) * length - u * u


RANDOM_SET = [False]  # set by main(): the workload's pairs are unrelated (set R)
RAGGED_SET = [False]  # set by main(): |a| ~ U[0.75 len, 1.25 len]


def make_inputs(op, n, length, k, costs, seed):
    from triple_accel_b200 import synth
    if op == "lev_k" and RANDOM_SET[0]:
        return synth.random_pairs(n, length, seed=seed)
    if op == "lev_k" and RAGGED_SET[0]:
        return synth.ragged_mutated_pairs(n, length * 3 // 4, length * 5 // 4, k, seed=seed)
    if op == "hamming":
        return synth.hamming_pairs(n, length, seed=seed)
    if op == "exp":
        return synth.mutated_pairs(n, length, 4, seed=seed, exact_edits=True)
    if op == "search":  # (needle, needle offsets placeholder, haystack bytes, haystack offsets)
        needle, hay, hoff = synth.needle_haystacks(n, length, 32, plant_frac=0.01, max_edits=3, seed=seed)
        return needle, np.array([0, len(needle)], np.uint64), hay, hoff
    return synth.mutated_pairs(n, length, k, seed=seed, allow_swap=bool(costs[3]))


def cells_per_unit(op, length, k):
    if op == "hamming":
        return length
    if op == "search":
        return 32 * length  # needle_len x haystack_len cells of the reference's scalar search DP
    return nominal_cells(length, k)


def oracle_run(orc, op, a, ao, b, bo, k, costs, cnt, threads):
    """the oracle on the first cnt units; returns something comparable with gpu_result()"""
    if op == "hamming":
        return orc.hamming_batch(a, ao[:cnt + 1], b, bo[:cnt + 1], threads=threads)
    if op == "exp":
        return orc.levenshtein_exp_batch(a, ao[:cnt + 1], b, bo[:cnt + 1], costs, threads=threads)
    if op == "search":
        m, off = orc.levenshtein_search_batch(a, b, bo[:cnt + 1], k, 1, costs, False, threads=threads)
        return np.concatenate([off.astype(np.uint64), m.reshape(-1)])
    return orc.levenshtein_k_batch(a, ao[:cnt + 1], b, bo[:cnt + 1], k, costs, threads=threads)


def reference_cpu_run(orc, op, a, ao, b, bo, k, costs, cnt, threads, length):
    """What a user of the reference runs on this host's CPU for this path.  On an AVX2 host the crate's public
    functions dispatch to its SIMD cores; oracle/ta_ref_avx2.c restates the one every BASELINE config selects
    (Avx1x32x8, and Avx::count_mismatches for Hamming), so that is what is timed when it covers the workload;
    otherwise (search, wide bands, no AVX2) the scalar port.  Returns (result, description)."""
    simd = orc.simd_available()
    if simd and op == "hamming":
        return orc.hamming_simd_batch(a, ao[:cnt + 1], b, bo[:cnt + 1], threads=threads), \
            "AVX2 restatement of hamming_simd_parallel (Avx::count_mismatches)"
    if simd and op == "exp":
        return orc.levenshtein_simd_exp_batch(a, ao[:cnt + 1], b, bo[:cnt + 1], costs, threads=threads), \
            "AVX2 restatement of levenshtein_exp over levenshtein_simd_k_with_opts (Avx1x32x8 core)"
    if simd and op == "lev_k" and orc.lib().orc_simd_covers(length, length, k, orc.Costs(*costs)):
        return orc.levenshtein_simd_k_batch(a, ao[:cnt + 1], b, bo[:cnt + 1], k, costs, threads=threads), \
            "AVX2 restatement of levenshtein_simd_k_with_opts (Avx1x32x8 core)"
    if simd and op == "search" and len(a) <= 32 and max(len(a) + k, k + 1) <= 255:
        m, off = orc.levenshtein_search_simd_batch(a, b, bo[:cnt + 1], k, 1, costs, False, threads=threads)
        return np.concatenate([off.astype(np.uint64), m.reshape(-1)]), \
            "AVX2 restatement of levenshtein_search_simd_with_opts (Avx1x32x8 search core)"
    return oracle_run(orc, op, a, ao, b, bo, k, costs, cnt, threads), "scalar oracle (port of the reference's scalar routine)"


def dominant_kernel(op, k, costs, length):
    """name of the kernel the dispatcher picks for this workload (triple_accel_b200/csrc/lev_bitpar.cu, search.cu)"""
    if op == "hamming":
        return "hamming_kernel"
    if op == "search":
        return "search_pigeon_staged_kernel (+ search_wave_kernel on the flagged 128-byte sub-segments)"
    unit = tuple(costs[:3]) == (1, 1, 0) and costs[3] <= 1
    if op == "exp":
        k = 15 if costs[3] else 16  # first round of the exponential search
    band = min(k, length) + 1 + (1 if costs[3] else 0)
    if not unit or band > 64:
        return "lev_band_kernel"
    if band <= 9 and not costs[3]:
        return "lev_bitpar_duo_kernel"
    if band <= 25:
        return "lev_bitpar_blk_kernel<C=%d>" % (16 if band <= 17 else 8)
    return "lev_bitpar_tab_kernel<%s>" % ("u32" if band <= 32 else "u64")


class ClockSampler:
    """samples nvidia-smi SM clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def pinned_copy(lib, arr):
    """copy a numpy array into pinned host memory obtained from the C ABI"""
    import ctypes as C
    nbytes = max(arr.nbytes, 1)
    p = lib.ta_host_alloc(nbytes)
    if not p:
        raise MemoryError("ta_host_alloc")
    buf = (C.c_uint8 * nbytes).from_address(p)
    out = np.frombuffer(buf, dtype=arr.dtype, count=arr.size)
    out[:] = arr.reshape(-1)
    return out, p


def run_reference(args, wl):
    """--impl reference: the reference's own CPU algorithm for this path.  The crate is Rust-only and cannot be
    built in this image, so this times a C restatement (kind "port") with every host thread, on a bounded sample:
    the crate's AVX2 code path where oracle/ta_ref_avx2.c covers the workload, else its scalar routine."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as orc
    op, n, length, k, costs, desc = wl
    if op == "lev_k_R":
        op, RANDOM_SET[0] = "lev_k", True
    if op == "lev_k_ragged":
        op, RAGGED_SET[0] = "lev_k", True
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the reference arm works on a larger bounded sample than the in-line cpu_baseline (it has the run to itself)
    ref_sample = max(args.cpu_sample, 1_000_000)
    sample = min(n, ref_sample if op != "search" else max(1, ref_sample // 100))
    if op == "exp":
        sample = min(sample, 100_000)
    a, ao, b, bo = make_inputs(op, sample, length, k, costs, 1234)
    threads = orc.max_threads()

    what = [""]

    def step():
        r, what[0] = reference_cpu_run(orc, op, a, ao, b, bo, k, costs, sample, threads, length)
        return r

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    cells = cells_per_unit(op, length, k)
    val = sample * cells / dt / 1e9
    line = {
        "impl": "reference", "metric": "dp_cell_updates_per_s", "value": val, "unit": "GCUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": desc, "name": args.workload, "sample_pairs": sample},
        "pairs_per_s": sample / dt,
        "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": threads, "kind": "port",
                         "sample": "%d units of the same workload per step; %s on %d threads"
                                   % (sample, what[0], threads)},
        "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lev_k8_len128", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="override the number of pairs per GPU")
    ap.add_argument("--cpu-sample", type=int, default=200_000, help="pairs per step for CPU baselines")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    wl = list(WORKLOADS[args.workload])
    if args.pairs:
        wl[1] = args.pairs
    if args.impl == "reference":
        run_reference(args, wl)
        return
    op, n, length, k, costs, desc = wl
    if op == "lev_k_R":
        op, RANDOM_SET[0] = "lev_k", True
    if op == "lev_k_ragged":
        op, RAGGED_SET[0] = "lev_k", True

    import torch
    import triple_accel_b200 as ta
    from triple_accel_b200 import _ffi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    eng = ta.Engine(local_rank)
    lib = _ffi.load()

    # every rank owns an independent shard of the same shape (weak scaling; no data-path collective)
    a, ao, b, bo = make_inputs(op, n, length, k, costs, 1234 + rank)
    max_len = int(max((ao[1:] - ao[:-1]).max(), (bo[1:] - bo[:-1]).max())) if n else 0

    def to_dev(x):
        return torch.from_numpy(x.view(np.int64) if x.dtype == np.uint64 else x).to(dev)

    d_a, d_ao, d_b, d_bo = to_dev(a), to_dev(ao), to_dev(b), to_dev(bo)
    d_out = torch.empty(n, dtype=torch.int32, device=dev)

    last = {}

    def step_dev():
        if op == "hamming":
            eng.hamming_batch_dev(d_a, d_ao, d_b, d_bo, d_out)
        elif op == "exp":
            eng.levenshtein_exp_batch_dev(d_a, d_ao, d_b, d_bo, costs, max_len, d_out)
        elif op == "search":
            last["m"] = eng.levenshtein_search_batch_dev(a, d_b, d_bo, max_len, k, 1, costs, False)
        else:
            eng.levenshtein_k_batch_dev(d_a, d_ao, d_b, d_bo, k, costs, max_len, d_out)

    def gpu_result(cnt):
        if op == "search":
            m, off = last["m"]
            tot = int(off[cnt])
            return np.concatenate([off[:cnt + 1].astype(np.uint64), m[:tot].reshape(-1)])
        return d_out[:cnt].cpu().numpy().view(np.uint32)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_dev()
    eng.dev_status()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_dev()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    eng.dev_status()
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / args.steps

    # parity spot check on the timed output (bounded sample, oracle as the checker)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as orc
    chk = min(n, {"search": 2000, "exp": 5000}.get(op, 20000))
    got = gpu_result(chk)
    want = oracle_run(orc, op, a, ao, b, bo, k, costs, chk, orc.max_threads())
    parity_ok = bool(np.array_equal(got, want))

    cells_pair = cells_per_unit(op, length, k)
    total_pairs = n * world
    value = total_pairs * cells_pair / (ms_step * 1e-3) / 1e9
    # algorithmic bytes: |a| + |b| + 4 per pair (CSR offsets, +16 B/pair, not counted); search: |haystack|
    alg_bytes = int(b.nbytes) if op == "search" else int(a.nbytes + b.nbytes + 4 * n)
    if RANDOM_SET[0]:
        # set R: a pair is decided at the first 16-column boundary past k + 1 columns; count only the bytes needed to
        # get there ("if the kernel early-exits, only count bytes of strings it actually streamed" --
        # the kernel's prefetches read somewhat more, so this is the conservative figure)
        need = min(length, 16 * ((k + 1 + 15) // 16))
        alg_bytes = int(n * (2 * need + 4))

    # ---- end to end through the host-buffer C ABI (pinned host inputs; H2D + kernel + D2H timed) --------------
    e2e = None
    if not args.no_e2e:
        pa, _p1 = pinned_copy(lib, a)
        pb, _p2 = pinned_copy(lib, b)
        pao, _p3 = pinned_copy(lib, ao)
        pbo, _p4 = pinned_copy(lib, bo)
        pout, _p5 = pinned_copy(lib, np.zeros(n, np.uint32))

        def step_host():
            if op == "hamming":
                eng.hamming_batch(pa, pao, pb, pbo, out=pout)
            elif op == "exp":
                eng.levenshtein_exp_batch(pa, pao, pb, pbo, costs, out=pout)
            elif op == "search":
                last["m"] = eng.levenshtein_search_batch(pa, pb, pbo, k, 1, costs, False)
            else:
                eng.levenshtein_k_batch(pa, pao, pb, pbo, k, costs, out=pout)

        for _ in range(2):
            step_host()
        barrier()
        e_steps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(e_steps):
            step_host()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        dt /= e_steps
        assert np.array_equal(gpu_result(chk) if op == "search" else pout[:chk], want), "e2e parity"
        e2e = {"value": total_pairs * cells_pair / dt / 1e9, "unit": "GCUPS",
               "h2d_bytes_per_step": int(a.nbytes + b.nbytes + ao.nbytes + bo.nbytes), "d2h_bytes_per_step": int(4 * n),
               "ms_per_step": dt * 1e3, "pairs_per_s": total_pairs / dt}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (ms_step * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "kernel": dominant_kernel(op, k, costs, length)}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        sample = min(n, args.cpu_sample if op != "search" else max(1, args.cpu_sample // 100))
        if op == "exp":
            sample = min(sample, 20000)
        threads = orc.max_threads()
        t0 = time.perf_counter()
        _, what = reference_cpu_run(orc, op, a, ao, b, bo, k, costs, sample, threads, length)
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        oracle_run(orc, op, a, ao, b, bo, k, costs, sample, threads)
        dt_scalar = time.perf_counter() - t0
        one = max(1, sample // threads)  # the same reference path on ONE thread (the crate itself is single-threaded)
        t0 = time.perf_counter()
        reference_cpu_run(orc, op, a, ao, b, bo, k, costs, one, 1, length)
        dt_one = time.perf_counter() - t0
        cpu_baseline = {"value": sample * cells_pair / dt / 1e9, "unit": "GCUPS", "cores": threads, "kind": "port",
                        "sample": "first %d units of the same batch; %s on %d threads" % (sample, what, threads),
                        "pairs_per_s": sample / dt, "scalar_port_pairs_per_s": sample / dt_scalar,
                        "one_thread_pairs_per_s": one / dt_one}

    line = {
        "metric": "dp_cell_updates_per_s", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": desc, "name": args.workload, "pairs_per_gpu": n, "len": length, "k": k,
                   "costs": list(costs), "cells_per_pair": cells_pair, "parallelism": "pairs sharded x%d" % world,
                   "l2": "inputs (%.0f MB per GPU) larger than the 126 MB L2" % ((a.nbytes + b.nbytes) / 1e6)
                   if a.nbytes + b.nbytes > 126e6 else "inputs fit in L2 (small config)"},
        "pairs_per_s": total_pairs / (ms_step * 1e-3),
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "cpu_baseline": cpu_baseline, "parity_checked_pairs": chk, "parity_ok": parity_ok,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    if not parity_ok:
        raise SystemExit("bench.py: GPU results differ from the oracle")


if __name__ == "__main__":
    main()
