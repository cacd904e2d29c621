"""The search path's host phase (hit ordering, row-0 match, running Best threshold, Best post-pass:
triple_accel_b200/csrc/search_emit.hpp) compiled on its own and pinned to the scalar oracle on the CPU -- the oracle's
SearchType::All output, shuffled, stands in for the device phase (tests/cpp/search_emit_host.cpp)."""
import os
import shutil
import subprocess

import pytest

import _oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [1, 2])
def test_emit_rules_match_the_oracle(tmp_path, seed):
    if not shutil.which("g++"):
        pytest.skip("no g++ on this box")
    orc.build()
    exe = str(tmp_path / "search_emit_host")
    odir = os.path.join(ROOT, "oracle")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-w", os.path.join(ROOT, "tests", "cpp", "search_emit_host.cpp"),
                           "-L", odir, "-lta_oracle", "-Wl,-rpath," + odir, "-o", exe])
    r = subprocess.run([exe, "300", str(seed)], capture_output=True, text=True)
    assert r.returncode == 0 and " bad 0" in r.stdout, r.stdout[-2000:] + r.stderr[-1000:]
