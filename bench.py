#!/usr/bin/env python3
"""bench.py -- throughput of the hot path on B200 (see DESIGN.md "Measurement").

A "step" is one pass of the hot path over one batch of synthetic input.  The headline (`value`, `e2e`, `roofline`,
`cpu_baseline`) is BASELINE.json configs[1]: levenshtein_simd_k, k = 8, 1 M pairs, len 128, unit costs, on the
"matching" set M (b = a after U[0,8] random edits, so no pair can be rejected early).  The same line
carries a `configs` array with every other BASELINE config and the north-star lines (k = 16 at len 128 and len 4096,
cfg 1, 3, 4, 5, the unrelated set R, ragged lengths, a general cost model), each measured the same way with fewer
steps; cfg 4 and cfg 5 -- the configs BASELINE.json shards over 8 GPUs -- are ONE batch split across the ranks
("scaling": "strong"), the rest replicate the per-GPU batch ("weak").

  python bench.py --gpus 1 --steps 20 --warmup 5            # our CUDA path (one JSON line)
  python bench.py --impl reference --steps 3 --warmup 1     # the reference's CPU algorithm (C restatement) on host cores
  torchrun --nproc-per-node N bench.py --gpus N ...          # one rank per GPU, no data-path collective
  python bench.py --gpus N --inproc --workload W             # ONE process, ONE multi-device context (ta_init_multi)

value  = GCUPS (nominal band cells of the reference's scalar banded DP, (2k+1)n - k^2 per pair) with inputs
         resident in HBM, timed with CUDA events on the launching stream, max over ranks.
e2e    = the same metric through the host-buffer C-ABI call (pinned host inputs, H2D + kernel + D2H in the timed
         region), next to the bare concurrent pinned-copy bandwidth of the same ranks (`staging`).
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
    # name: (op, units per GPU (weak) or in total (strong), len, k, costs, scaling, description)
    "lev_k8_len128": ("lev_k", 1_000_000, 128, 8, (1, 1, 0, 0), "weak",
                      "levenshtein_simd_k k=8, 1M pairs len=128, unit costs, set M (b = a after U[0,8] edits)"),
    "lev_k16_len128": ("lev_k", 1_000_000, 128, 16, (1, 1, 0, 0), "weak",
                       "levenshtein_simd_k k=16, 1M pairs len=128, unit costs, set M (b = a after U[0,16] edits)"),
    "lev_k8_len128_R": ("lev_k_R", 1_000_000, 128, 8, (1, 1, 0, 0), "weak",
                        "levenshtein_simd_k k=8, 1M pairs len=128, unit costs, set R (b independent of a: every pair is "
                        "None; the kernel leaves a pair once its final-diagonal value exceeds k)"),
    "lev_k16_len128_R": ("lev_k_R", 1_000_000, 128, 16, (1, 1, 0, 0), "weak",
                         "levenshtein_simd_k k=16, 1M pairs len=128, unit costs, set R (b independent of a)"),
    "lev_k8_ragged96_160": ("lev_k_ragged", 1_000_000, 128, 8, (1, 1, 0, 0), "weak",
                            "levenshtein_simd_k k=8, 1M pairs, |a| ~ U[96,160] (mean 128), unit costs, set M: neighbouring "
                            "pairs have unrelated lengths (arbitrary alignment, ragged tails)"),
    "lev_k16_ragged96_160": ("lev_k_ragged", 1_000_000, 128, 16, (1, 1, 0, 0), "weak",
                             "levenshtein_simd_k k=16, 1M pairs, |a| ~ U[96,160], unit costs, set M (one pair per thread)"),
    "rdamerau_k16_len512": ("lev_k", 1_000_000, 512, 16, (1, 1, 0, 1), "strong",
                            "RDAMERAU_COSTS k=16, 1M pairs len=512, set M (U[0,16] edits incl. swaps), one batch over all GPUs"),
    "lev_k16_len4096": ("lev_k", 262_144, 4096, 16, (1, 1, 0, 0), "weak",
                        "levenshtein_simd_k k=16, 256Ki pairs len=4096, unit costs, set M"),
    "affine_k16_len128": ("lev_k", 1_000_000, 128, 16, (2, 1, 3, 0), "weak",
                          "levenshtein_simd_k_with_opts k=16, 1M pairs len=128, EditCosts(2,1,3,None) (general kernel)"),
    "lev_k60_len1024": ("lev_k", 262_144, 1024, 60, (1, 1, 0, 0), "weak",
                        "levenshtein_simd_k k=60, 256Ki pairs len=1024, unit costs, set M (64-row window kernel)"),
    "exp_len1024": ("exp", 1_000_000, 1024, 30, (1, 1, 0, 0), "weak",
                    "levenshtein_exp, 1M pairs len=1024, exactly 4 random edits (first round succeeds)"),
    "search_n32_h4096": ("search", 100_000, 4096, 3, (1, 1, 0, 0), "strong",
                         "levenshtein_search needle len=32 over 100k haystacks len=4096, k=3, Best, 1% planted hits, one "
                         "batch over all GPUs, needle broadcast from rank 0 with NCCL"),
    "search_all_n32_h4096": ("search", 100_000, 4096, 3, (1, 1, 0, 0), "strong",
                             "levenshtein_search SearchType::All, needle len=32, 100k haystacks len=4096, k=3, 1% planted hits"),
    "search_n64_h4096": ("search", 100_000, 4096, 6, (1, 1, 0, 0), "strong",
                         "levenshtein_search needle len=64 (seven pieces of 9), 100k haystacks len=4096, k=6, Best, 1% planted hits"),
    "search_affine_n32_h4096": ("search", 20_000, 4096, 6, (2, 1, 3, 0), "strong",
                                "levenshtein_search EditCosts(2,1,3,None): unit-cost pre-filter with the operation-count bound, exact (cost, length) kernel on the flagged "
                                "haystacks; needle len=32, 20k haystacks len=4096, k=6, Best"),
    "hamming_len64": ("hamming", 10_000, 64, 0, (1, 1, 0, 0), "weak", "hamming, 10k pairs len=64 (plumbing case)"),
    "hamming_len4096": ("hamming", 262_144, 4096, 0, (1, 1, 0, 0), "weak", "hamming, 256Ki pairs len=4096"),
}
SEARCH_OPTS = {  # needle length, SearchType (0 All, 1 Best) of the search workloads
    "search_n32_h4096": (32, 1), "search_all_n32_h4096": (32, 0), "search_n64_h4096": (64, 1),
    "search_affine_n32_h4096": (32, 1),
}
HEADLINE = "lev_k8_len128"
# the other BASELINE configs and north-star lines, reported in the `configs` array of the default line
CONFIG_ARRAY = ["lev_k16_len128", "lev_k16_len4096", "hamming_len64", "exp_len1024", "search_n32_h4096",
                "rdamerau_k16_len512", "lev_k8_len128_R", "lev_k8_ragged96_160", "lev_k16_ragged96_160", "affine_k16_len128",
                "search_all_n32_h4096", "search_n64_h4096", "search_affine_n32_h4096"]
REF_THREADS = 16  # the reference arm and cpu_baseline use min(16, host threads): comparable between boxes


def nominal_cells(length, k):
    """cells the reference's scalar banded DP visits for |a| = |b| = length, half-width k"""
    u = min(k, length)
    return (2 * u + 1) * length - u * u


def base_op(op):
    return "lev_k" if op.startswith("lev_k") else op


def make_inputs(op, n, length, k, costs, seed, first=0, needle=None):
    """units first .. first + n of the workload's batch (every unit has its own random stream: tools/ta_synth.c)"""
    from triple_accel_b200 import synth
    if op == "lev_k_R":
        return synth.random_pairs(n, length, seed=seed + first)
    if op == "lev_k_ragged":
        return synth.edited_pairs(n, length * 3 // 4, length * 5 // 4, k, seed=seed, first=first)
    if op == "hamming":
        return synth.hamming_pairs(n, length, seed=seed + first)
    if op == "exp":
        return synth.edited_pairs(n, length, length, 4, seed=seed, first=first, exact_edits=True)
    if op == "search":  # (needle, needle offsets placeholder, haystack bytes, haystack offsets)
        hay, hoff = synth.planted_haystacks(n, length, needle, plant_frac=0.01, max_edits=3, seed=seed, first=first)
        return needle, np.array([0, len(needle)], np.uint64), hay, hoff
    return synth.edited_pairs(n, length, length, k, seed=seed, first=first, allow_swap=bool(costs[3]))


def make_needle(seed, length=32):
    return np.random.Generator(np.random.PCG64(seed)).integers(1, 256, size=length, dtype=np.uint8)


def cells_per_unit(op, length, k, needle_len=32):
    if op == "hamming":
        return length
    if op == "search":
        return needle_len * length  # needle_len x haystack_len cells of the reference's scalar search DP
    return nominal_cells(length, k)


SEARCH_TYPE = [1]  # SearchType of the search workload being measured (set by the callers from SEARCH_OPTS)
SEARCH_NEEDLE = [32]  # its needle length


def oracle_run(orc, op, a, ao, b, bo, k, costs, cnt, threads):
    """the oracle on the first cnt units; returns something comparable with gpu_result()"""
    if op == "hamming":
        return orc.hamming_batch(a, ao[:cnt + 1], b, bo[:cnt + 1], threads=threads)
    if op == "exp":
        return orc.levenshtein_exp_batch(a, ao[:cnt + 1], b, bo[:cnt + 1], costs, threads=threads)
    if op == "search":
        m, off = orc.levenshtein_search_batch(a, b, bo[:cnt + 1], k, SEARCH_TYPE[0], costs, False, threads=threads)
        return np.concatenate([off.astype(np.uint64), m.reshape(-1)])
    return orc.levenshtein_k_batch(a, ao[:cnt + 1], b, bo[:cnt + 1], k, costs, threads=threads)


def reference_cpu_run(orc, op, a, ao, b, bo, k, costs, cnt, threads, length):
    """What a user of the reference runs on this host's CPU for this path.  On an AVX2 host the crate's public
    functions dispatch to its SIMD cores; oracle/ta_ref_avx2.c restates the one every BASELINE config selects
    (Avx1x32x8, and Avx::count_mismatches for Hamming), so that is what is timed when it covers the workload;
    otherwise (wide bands, no AVX2) the scalar port.  Returns (result, description)."""
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
    if simd and op == "search" and len(a) <= 32 and max(len(a) + k, k + 1) <= 255 and tuple(costs) == (1, 1, 0, 0):
        m, off = orc.levenshtein_search_simd_batch(a, b, bo[:cnt + 1], k, SEARCH_TYPE[0], costs, False, threads=threads)
        return np.concatenate([off.astype(np.uint64), m.reshape(-1)]), \
            "AVX2 restatement of levenshtein_search_simd_with_opts (Avx1x32x8 search core)"
    return oracle_run(orc, op, a, ao, b, bo, k, costs, cnt, threads), "scalar oracle (port of the reference's scalar routine)"


def dominant_kernel(op, k, costs, length, n_units=None, ragged=False):
    """name of the kernel the dispatcher picks for this workload (triple_accel_b200/csrc/api.cu: ta_launch_lev)"""
    if op == "hamming":
        return "hamming_kernel"
    unit = tuple(costs[:3]) == (1, 1, 0) and costs[3] <= 1
    if op == "search":
        nlen = SEARCH_NEEDLE[0]
        # weighted costs run the unit-cost pre-filters with the number of edit operations a match can hold (search.cu)
        ku = k if unit else k // min(c for c in (costs[0], costs[1], costs[3]) if c)
        if ku >= nlen or nlen > 64:
            return "search_exact_kernel / search_wave_kernel on whole haystacks (no pre-filter)"
        k = ku
        pieces = (2 * k + 1) if costs[3] else (k + 1)
        forced = os.environ.get("TA_SEARCH_FILTER", "")
        big = n_units is None or n_units * length >= ((64 << 20) if nlen <= 32 else (8 << 20))  # lev_bitpar.cu dispatch
        if nlen <= 64 and pieces <= nlen and nlen // pieces >= 7 and (forced == "qgram" or (forced == "" and big)):
            return "search_qgram_kernel (+ search_qgram_resolve_kernel, search_wave_kernel on the flagged 16-byte granules)"
        if nlen <= 32 and pieces <= nlen and nlen // pieces >= 4 and forced != "myers":
            return "search_pigeon_staged_kernel (+ search_wave_kernel on the flagged 128-byte sub-segments)"
        return "search_filter_kernel (+ search_wave_kernel on the flagged 128-byte sub-segments)"
    if op == "exp":
        k = 15 if costs[3] else 16  # first round of the exponential search
    kk = min(k, length)
    if unit and kk <= 64 and length >= 1024 and length >= 4 * kk * kk and os.environ.get("TA_FR", "") != "0":
        return "lev_fr_kernel"
    band = kk + 1 + (1 if costs[3] else 0)
    if not unit or band > 64:
        w_bound = (min(k, length * max(costs[0], costs[1]) + costs[2]) - costs[2]) // costs[1] + 1 + (2 if costs[3] else 0)
        return "lev_diag16_kernel" if w_bound <= 32 else "lev_band_kernel"
    if band <= 9 and not costs[3]:
        tiled = {"0": False, "1": True}.get(os.environ.get("TA_DUO_TILED", ""), ragged)
        return "lev_bitpar_duo_tiled_kernel" if tiled else "lev_bitpar_duo_kernel"
    if band <= 25:
        tiled = {"0": False, "1": True}.get(os.environ.get("TA_BLK_TILED", ""), ragged)
        return "lev_bitpar_blk%s_kernel<C=%d>" % ("_tiled" if tiled else "", 16 if band <= 17 else 8)
    return "lev_bitpar_tab_kernel<%s>" % ("u32" if band <= 32 else "u64")


def workload_config(name, world, units_per_gpu=None):
    """the `config` object of a line: identical in our arm and the reference arm"""
    op, n, length, k, costs, scaling, desc = WORKLOADS[name]
    n = units_per_gpu or n
    per_gpu_bytes = (n if scaling == "weak" else n // world) * length * (1 if base_op(op) == "search" else 2)
    return {"workload": desc, "name": name,
            "l2": "inputs (%.0f MB per GPU) larger than the 126 MB L2" % (per_gpu_bytes / 1e6) if per_gpu_bytes > 126e6
            else "inputs (%.1f MB per GPU) fit in L2 (small config)" % (per_gpu_bytes / 1e6), "units_per_gpu": n if scaling == "weak" else None,
            "units_total": n * world if scaling == "weak" else n, "len": length, "k": k, "costs": list(costs),
            "cells_per_unit": cells_per_unit(base_op(op), length, k, SEARCH_OPTS.get(name, (32, 1))[0]), "scaling": scaling,
            "parallelism": "units sharded x%d" % world}


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


class Pinned:
    """numpy views of pinned host memory obtained from the C ABI (ta_host_alloc), released on close()"""

    def __init__(self, lib):
        self.lib, self.ptrs = lib, []

    def copy(self, arr):
        import ctypes as C
        nbytes = max(arr.nbytes, 1)
        p = self.lib.ta_host_alloc(nbytes)
        if not p:
            raise MemoryError("ta_host_alloc")
        self.ptrs.append(p)
        out = np.frombuffer((C.c_uint8 * nbytes).from_address(p), dtype=arr.dtype, count=arr.size)
        out[:] = arr.reshape(-1)
        return out

    def close(self):
        for p in self.ptrs:
            self.lib.ta_host_free(p)
        self.ptrs = []


def ref_threads(orc):
    return max(1, min(REF_THREADS, orc.max_threads()))


def run_reference(args, name):
    """--impl reference: the reference's own CPU algorithm for this path.  The crate is Rust-only and neither this
    image nor the GPU box has rustc/cargo (profiles/r02_probe_gpu_box.txt), so this times a C restatement (kind
    "port") on min(16, host threads) threads, on a bounded sample: the crate's AVX2 code path where
    oracle/ta_ref_avx2.c covers the workload, else its scalar routine."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as orc
    op, n, length, k, costs, scaling, desc = WORKLOADS[name]
    if args.pairs:
        n = args.pairs
    if int(os.environ.get("RANK", "0")) != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # the reference arm works on a larger bounded sample than the in-line cpu_baseline (it has the run to itself)
    ref_sample = max(args.cpu_sample, 1_000_000)
    sample = min(n, ref_sample if op != "search" else max(1, ref_sample // 100))
    if op == "exp":
        sample = min(sample, 100_000)
    nlen, SEARCH_TYPE[0] = SEARCH_OPTS.get(name, (32, 1))
    SEARCH_NEEDLE[0] = nlen
    needle = make_needle(1234, nlen) if op == "search" else None
    a, ao, b, bo = make_inputs(op, sample, length, k, costs, 1234, needle=needle)
    threads = ref_threads(orc)
    bop = base_op(op)
    what = [""]

    def step():
        r, what[0] = reference_cpu_run(orc, bop, a, ao, b, bo, k, costs, sample, threads, length)
        return r

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    cells = cells_per_unit(bop, length, k, nlen)
    val = sample * cells / dt / 1e9
    line = {
        "impl": "reference", "metric": "dp_cell_updates_per_s", "value": val, "unit": "GCUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(name, world, args.pairs or None),
        "pairs_per_s": sample / dt,
        "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": threads, "kind": "port",
                         "sample": "%d units of the same workload per step; %s on %d threads (host has %d)"
                                   % (sample, what[0], threads, orc.max_threads())},
        "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


class Runner:
    """measures workloads on this rank's GPU; world > 1: one rank per GPU, max over ranks"""

    def __init__(self, args):
        import torch
        import triple_accel_b200 as ta
        from triple_accel_b200 import _ffi
        self.torch, self.ta, self.args = torch, ta, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- this framework has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.eng = ta.Engine(self.local_rank)
        self.lib = _ffi.load()
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import _oracle as orc
        self.orc = orc
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            self.peak, self.peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            self.peak, self.peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        cpath = os.path.join(ROOT, "profiles", "ncu_counts.json")
        self.ncu_counts = json.load(open(cpath)) if os.path.exists(cpath) else {}
        self.sm_count = torch.cuda.get_device_properties(self.dev).multi_processor_count

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([float(x)], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([float(x)], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def staging_ceiling(self, mb=256, reps=4):
        """bare pinned host-to-device copies, all ranks at once: what the box can feed its GPUs (GB/s per GPU)"""
        torch = self.torch
        h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
        d = torch.empty(mb << 20, dtype=torch.uint8, device=self.dev)
        d.copy_(h, non_blocking=True)
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dt = self.max_over_ranks(dt)
        del h, d
        return (mb << 20) * reps / dt / 1e9

    def measure(self, name, steps, warmup, e2e=True, cpu_baseline=False, clocks=False, units=None):
        torch, eng, orc, args = self.torch, self.eng, self.orc, self.args
        op, n_spec, length, k, costs, scaling, desc = WORKLOADS[name]
        n_spec = units or n_spec
        bop = base_op(op)
        world, rank = self.world, self.rank
        # weak: every rank owns an independent batch of the same shape; strong: rank r owns units [lo, hi) of ONE batch
        if scaling == "strong" and world > 1:
            lo, hi = n_spec * rank // world, n_spec * (rank + 1) // world
            n, first, total_units = hi - lo, lo, n_spec
        else:
            n, first, total_units = n_spec, n_spec * rank, n_spec * world
            if scaling == "strong":
                first = 0
        needle = None
        nlen, stype = SEARCH_OPTS.get(name, (32, 1))
        SEARCH_TYPE[0] = stype
        SEARCH_NEEDLE[0] = nlen
        if op == "search":
            # rank 0 alone knows the needle; the others receive it by NCCL broadcast (torch.distributed)
            from triple_accel_b200 import dist as tdist
            mine = make_needle(1234, nlen) if rank == 0 else np.zeros(1, np.uint8)
            needle = tdist.broadcast_needle(mine) if self.dist is not None else mine
        a, ao, b, bo = make_inputs(op, n, length, k, costs, 1234, first=first, needle=needle)
        max_len = int(max((ao[1:] - ao[:-1]).max(), (bo[1:] - bo[:-1]).max())) if n else 0

        def to_dev(x):
            return torch.from_numpy(x.view(np.int64) if x.dtype == np.uint64 else x).to(self.dev)

        d_a, d_ao, d_b, d_bo = to_dev(a), to_dev(ao), to_dev(b), to_dev(bo)
        d_out = torch.empty(max(n, 1), dtype=torch.int32, device=self.dev)
        last = {}
        # the device-resident entry points cannot look at the offsets: the caller says whether the lengths vary
        # (ta_set_length_hint; the host-buffer calls of the e2e leg decide from the offsets themselves)
        eng.set_length_hint(True if op == "lev_k_ragged" else None)

        def step_dev():
            if bop == "hamming":
                eng.hamming_batch_dev(d_a, d_ao, d_b, d_bo, d_out, mean_len=length)
            elif bop == "exp":
                eng.levenshtein_exp_batch_dev(d_a, d_ao, d_b, d_bo, costs, max_len, d_out)
            elif bop == "search":
                last["m"] = eng.levenshtein_search_batch_dev(a, d_b, d_bo, max_len, k, stype, costs, False)
            else:
                eng.levenshtein_k_batch_dev(d_a, d_ao, d_b, d_bo, k, costs, max_len, d_out)

        def gpu_result(cnt):
            if bop == "search":
                m, off = last["m"]
                tot = int(off[cnt])
                return np.concatenate([off[:cnt + 1].astype(np.uint64), m[:tot].reshape(-1)])
            return d_out[:cnt].cpu().numpy().view(np.uint32)

        for _ in range(max(warmup, 3)):
            step_dev()
        eng.dev_status()
        self.barrier()
        sampler = ClockSampler(self.local_rank) if clocks else None
        if sampler and rank == 0:
            sampler.start()
            time.sleep(0.15)
        launches0 = eng.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        ev0.record()
        for _ in range(steps):
            step_dev()
        ev1.record()
        self.barrier()
        ms = ev0.elapsed_time(ev1)
        launches = eng.launch_count - launches0
        clk = sampler.stop() if (sampler and rank == 0) else None
        eng.dev_status()
        ms_step = self.max_over_ranks(ms) / steps

        # parity spot check on the timed output (bounded sample, oracle as the checker)
        chk = min(n, {"search": 2000, "exp": 5000}.get(bop, 20000))
        if length >= 4096:
            chk = min(chk, 4000)
        got = gpu_result(chk)
        want = oracle_run(orc, bop, a, ao, b, bo, k, costs, chk, ref_threads(orc))
        parity_ok = bool(np.array_equal(got, want))

        cells_unit = cells_per_unit(bop, length, k, nlen)
        value = total_units * cells_unit / (ms_step * 1e-3) / 1e9
        # algorithmic bytes of this rank's launch: |a| + |b| + 4 per pair (CSR offsets, +16 B/pair, not counted);
        # search: |haystack|
        alg_bytes = int(b.nbytes) if bop == "search" else int(a.nbytes + b.nbytes + 4 * n)
        if op == "lev_k_R":
            # set R: a pair is decided at the first 16-column boundary past k + 1 columns; count only the bytes needed
            # to get there ("if the kernel early-exits, only count bytes of strings it actually streamed")
            need = min(length, 16 * ((k + 1 + 15) // 16))
            alg_bytes = int(n * (2 * need + 4))
        achieved = alg_bytes / (ms_step * 1e-3) / 1e9
        res = {
            "name": name, "scaling": scaling if world > 1 else "single-gpu", "ms_per_step": ms_step,
            "units_per_s": total_units / (ms_step * 1e-3), "value": value, "unit": "GCUPS", "steps": steps,
            "frac_hbm": achieved / self.peak, "achieved_gbs": achieved, "algorithmic_bytes_per_launch": alg_bytes,
            "kernel": dominant_kernel(bop, k, costs, length, n, op == "lev_k_ragged"), "gpu_launches": int(launches), "parity_ok": parity_ok,
            "parity_checked_units": chk,
        }
        cnt = self.ncu_counts.get(name)
        if cnt and cnt.get("warp_instructions"):
            # instruction-issue roofline: warp instructions of the dominant kernel (ncu, profiles/) over what the SMs can
            # issue in the measured time: 4 schedulers x 1 warp instruction per clock per SM
            clock_hz = (clk or {}).get("sm_mhz") or 1965.0
            peak_issue = self.sm_count * 4 * clock_hz * 1e6
            res["issue_roofline"] = {"warp_instructions_per_launch": cnt["warp_instructions"],
                                     "frac_of_issue_peak": cnt["warp_instructions"] * (n / cnt.get("units", n)) /
                                     (ms_step * 1e-3) / peak_issue,
                                     "source": cnt.get("source")}
            res["traffic"] = cnt.get("dram_bytes")
        out = {"res": res, "clocks": clk, "e2e": None, "cpu_baseline": None, "total_units": total_units, "n": n}

        # ---- end to end through the host-buffer C ABI (pinned host inputs; H2D + kernel + D2H timed) --------------
        if e2e:
            pin = Pinned(self.lib)
            pa, pb, pao, pbo = pin.copy(a), pin.copy(b), pin.copy(ao), pin.copy(bo)
            pout = pin.copy(np.zeros(max(n, 1), np.uint32))

            def step_host():
                if bop == "hamming":
                    eng.hamming_batch(pa, pao, pb, pbo, out=pout)
                elif bop == "exp":
                    eng.levenshtein_exp_batch(pa, pao, pb, pbo, costs, out=pout)
                elif bop == "search":
                    last["m"] = eng.levenshtein_search_batch(pa, pb, pbo, k, stype, costs, False)
                else:
                    eng.levenshtein_k_batch(pa, pao, pb, pbo, k, costs, out=pout)

            for _ in range(2):
                step_host()
            self.barrier()
            e_steps = max(3, min(steps, 10))
            t0 = time.perf_counter()
            for _ in range(e_steps):
                step_host()
            torch.cuda.synchronize()
            dt = self.max_over_ranks(time.perf_counter() - t0) / e_steps
            ok = np.array_equal(gpu_result(chk) if bop == "search" else pout[:chk], want)
            h2d = int(a.nbytes + b.nbytes + ao.nbytes + bo.nbytes) if bop != "search" else int(b.nbytes + bo.nbytes + a.nbytes)
            h2d_all = self.sum_over_ranks(h2d)
            out["e2e"] = {"value": total_units * cells_unit / dt / 1e9, "unit": "GCUPS",
                          "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(self.sum_over_ranks(4 * n)),
                          "ms_per_step": dt * 1e3, "units_per_s": total_units / dt,
                          "h2d_gbs": h2d_all / dt / 1e9, "parity_ok": bool(ok)}
            pin.close()
            if not ok:
                res["parity_ok"] = False

        if cpu_baseline and rank == 0:
            sample = min(n, args.cpu_sample if bop != "search" else max(1, args.cpu_sample // 100))
            if bop == "exp":
                sample = min(sample, 20000)
            threads = ref_threads(orc)
            t0 = time.perf_counter()
            _, what = reference_cpu_run(orc, bop, a, ao, b, bo, k, costs, sample, threads, length)
            dt = time.perf_counter() - t0
            t0 = time.perf_counter()
            oracle_run(orc, bop, a, ao, b, bo, k, costs, sample, threads)
            dt_scalar = time.perf_counter() - t0
            one = max(1, sample // threads)  # the same reference path on ONE thread (the crate itself is single-threaded)
            t0 = time.perf_counter()
            reference_cpu_run(orc, bop, a, ao, b, bo, k, costs, one, 1, length)
            dt_one = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": sample * cells_unit / dt / 1e9, "unit": "GCUPS", "cores": threads, "kind": "port",
                                   "sample": "first %d units of the same batch; %s on %d threads (host has %d)"
                                             % (sample, what, threads, orc.max_threads()),
                                   "pairs_per_s": sample / dt, "scalar_port_pairs_per_s": sample / dt_scalar,
                                   "one_thread_pairs_per_s": one / dt_one}
        del d_a, d_ao, d_b, d_bo, d_out
        return out


def run_inproc(args, name):
    """ONE process, ONE context over args.gpus devices (ta_init_multi): the host-buffer call splits the batch inside the
    library.  End to end only (the device-resident entry points belong to single-device contexts)."""
    import triple_accel_b200 as ta
    from triple_accel_b200 import _ffi
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as orc
    op, n, length, k, costs, scaling, desc = WORKLOADS[name]
    n = args.pairs or n
    bop = base_op(op)
    eng = ta.Engine(devices=list(range(args.gpus)))
    lib = _ffi.load()
    nlen, stype = SEARCH_OPTS.get(name, (32, 1))
    SEARCH_TYPE[0] = stype
    SEARCH_NEEDLE[0] = nlen
    needle = make_needle(1234, nlen) if op == "search" else None
    a, ao, b, bo = make_inputs(op, n, length, k, costs, 1234, needle=needle)
    pin = Pinned(lib)
    pa, pb, pao, pbo = pin.copy(a), pin.copy(b), pin.copy(ao), pin.copy(bo)
    pout = pin.copy(np.zeros(n, np.uint32))
    last = {}

    def step_host():
        if bop == "hamming":
            eng.hamming_batch(pa, pao, pb, pbo, out=pout)
        elif bop == "exp":
            eng.levenshtein_exp_batch(pa, pao, pb, pbo, costs, out=pout)
        elif bop == "search":
            last["m"] = eng.levenshtein_search_batch(pa, pb, pbo, k, stype, costs, False)
        else:
            eng.levenshtein_k_batch(pa, pao, pb, pbo, k, costs, out=pout)

    for _ in range(max(args.warmup, 3)):
        step_host()
    l0 = eng.launch_count
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    dt = (time.perf_counter() - t0) / args.steps
    chk = min(n, {"search": 2000, "exp": 5000}.get(bop, 20000))
    want = oracle_run(orc, bop, a, ao, b, bo, k, costs, chk, ref_threads(orc))
    if bop == "search":
        m, off = last["m"]
        got = np.concatenate([off[:chk + 1].astype(np.uint64), m[:int(off[chk])].reshape(-1)])
    else:
        got = pout[:chk]
    cells_unit = cells_per_unit(bop, length, k, nlen)
    h2d = int(a.nbytes + b.nbytes + ao.nbytes + bo.nbytes)
    line = {"metric": "dp_cell_updates_per_s", "mode": "inproc-multi (ta_init_multi, one host-buffer call per step)",
            "value": n * cells_unit / dt / 1e9, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "dtype": "u32", "data": "synthetic", "config": workload_config(name, 1, args.pairs or None),
            "units_per_s": n / dt, "gpu_launches": eng.launch_count - l0, "uses_nccl": eng.uses_nccl,
            "needle_broadcasts": eng.needle_broadcasts,
            "e2e": {"value": n * cells_unit / dt / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4 * n, "h2d_gbs": h2d / dt / 1e9},
            "parity_ok": bool(np.array_equal(got, want)), "parity_checked_units": chk}
    print(json.dumps(line))
    pin.close()
    eng.close()
    if not line["parity_ok"]:
        raise SystemExit("bench.py: GPU results differ from the oracle")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="override the number of units (per GPU if weak, in total if strong)")
    ap.add_argument("--cpu-sample", type=int, default=200_000, help="pairs per step for CPU baselines")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="headline only (skip the `configs` array)")
    ap.add_argument("--config-steps", type=int, default=20)
    ap.add_argument("--inproc", action="store_true", help="one process, one multi-device context over --gpus devices")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, args.workload)
        return
    if args.inproc:
        run_inproc(args, args.workload)
        return

    t_start = time.perf_counter()
    R = Runner(args)
    head = R.measure(args.workload, args.steps, args.warmup, e2e=not args.no_e2e,
                     cpu_baseline=not args.no_cpu_baseline, clocks=True, units=args.pairs or None)
    staging = None
    if not args.no_e2e:
        per_gpu = R.staging_ceiling()
        e = head["e2e"]
        staging = {"bare_pinned_h2d_gbs_per_gpu": per_gpu, "bare_pinned_h2d_gbs_all_gpus": per_gpu * R.world,
                   "e2e_h2d_gbs_all_gpus": e["h2d_gbs"], "e2e_frac_of_bare_copy": e["h2d_gbs"] / (per_gpu * R.world),
                   "how": "all ranks copy 256 MB of pinned host memory to their GPU at once (torch copy_, 4 reps, max "
                          "over ranks); e2e moves h2d_bytes_per_step in ms_per_step"}
    configs = []
    if args.workload == HEADLINE and not args.no_configs and not args.pairs:
        for name in CONFIG_ARRAY:
            if time.perf_counter() - t_start > 150:  # keep the default run within a few minutes whatever the box
                configs.append({"name": name, "skipped": "time budget"})
                continue
            c = R.measure(name, args.config_steps, 3, e2e=not args.no_e2e)
            entry = dict(c["res"])
            entry["config"] = workload_config(name, R.world)
            if c["e2e"]:
                entry["e2e_ms_per_step"] = c["e2e"]["ms_per_step"]
                entry["e2e_units_per_s"] = c["e2e"]["units_per_s"]
                entry["e2e_h2d_gbs"] = c["e2e"]["h2d_gbs"]
            configs.append(entry)

    if R.rank == 0:
        res = head["res"]
        op, n, length, k, costs, scaling, desc = WORKLOADS[args.workload]
        roofline = {"bound": "hbm", "achieved": res["achieved_gbs"], "peak": R.peak, "unit": "GB/s",
                    "frac": res["frac_hbm"], "traffic": res.get("traffic"), "peak_source": R.peak_src,
                    "algorithmic_bytes_per_launch": res["algorithmic_bytes_per_launch"], "kernel": res["kernel"],
                    "issue": res.get("issue_roofline")}
        cfg = workload_config(args.workload, R.world, args.pairs or None)
        line = {
            "metric": "dp_cell_updates_per_s", "value": res["value"], "unit": "GCUPS", "n_gpus": R.world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": scaling if R.world > 1 else "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": cfg, "pairs_per_s": res["units_per_s"], "e2e": head["e2e"],
            "gpu_launches": res["gpu_launches"], "clocks": head["clocks"], "roofline": roofline,
            "cpu_baseline": head["cpu_baseline"], "staging": staging,
            "parity_checked_pairs": res["parity_checked_units"], "parity_ok": res["parity_ok"], "configs": configs,
            "bench_wall_s": time.perf_counter() - t_start,
        }
        print(json.dumps(line))
    all_ok = head["res"]["parity_ok"] and all(c.get("parity_ok", True) for c in configs)
    if R.dist is not None:
        R.dist.destroy_process_group()
    if not all_ok:
        raise SystemExit("bench.py: GPU results differ from the oracle")


if __name__ == "__main__":
    main()
