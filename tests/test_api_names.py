"""Drop-in surface: every `pub fn` / public type of the reference crate (src/lib.rs:126-127, 134-174, 196-235;
src/hamming.rs, src/levenshtein.rs `pub fn`s) resolves under the same name in the host mirror, and bench.py's
reference arm (the CPU restatement the driver times beside the GPU arm) runs without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CRATE_FNS = [
    # src/lib.rs
    "alloc_str", "fill_str",
    # src/hamming.rs
    "hamming_naive", "hamming_search_naive", "hamming_search_naive_with_opts", "hamming_words_64", "hamming_words_128",
    "hamming_simd_parallel", "hamming_simd_movemask", "hamming", "hamming_search_simd", "hamming_search_simd_with_opts",
    "hamming_search",
    # src/levenshtein.rs
    "levenshtein_naive", "levenstein_naive_str", "levenshtein_naive_with_opts", "levenshtein_naive_k",
    "levenshtein_naive_k_with_opts", "levenshtein_simd_k_str", "levenshtein_simd_k", "levenshtein_simd_k_with_opts",
    "levenshtein", "rdamerau", "levenshtein_exp", "levenshtein_exp_with_opts", "rdamerau_exp",
    "levenshtein_search_naive", "levenshtein_search_naive_with_opts", "levenshtein_search_simd",
    "levenshtein_search_simd_with_opts", "levenshtein_search",
]
CRATE_TYPES = ["Match", "Edit", "EditType", "SearchType", "EditCosts", "LEVENSHTEIN_COSTS", "RDAMERAU_COSTS"]


def test_every_public_name_of_the_crate_resolves():
    import triple_accel_b200 as ta
    for name in CRATE_FNS + CRATE_TYPES:
        assert hasattr(ta, name), name
    for name in CRATE_FNS:
        if name not in ("alloc_str", "fill_str"):
            assert hasattr(ta.Engine, name), name
    buf = ta.alloc_str(19)
    assert len(buf) == 19 and not any(buf)
    ta.fill_str(buf, b"abc")
    assert bytes(buf[:4]) == b"abc\0"
    et = ta.EditType  # src/lib.rs:147-158, codes in declaration order
    assert (et.Match, et.Mismatch, et.AGap, et.BGap, et.Transpose) == (0, 1, 2, 3, 4)


def test_bench_reference_arm_runs_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--pairs", "20000"], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "dp_cell_updates_per_s" and line["unit"] == "GCUPS"
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
