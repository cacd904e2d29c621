"""The per-pair bit-parallel cores (triple_accel_b200/csrc/lev_bitpar_core.cuh is host/device code) compiled for the host
and pinned to the scalar oracle: every table variant (sliding 16/32/64-row, block table with 16- and 8-position
blocks, 128/256 entries, multiply-add shift forms), random alphabets of 2..256 symbols, ragged lengths, misaligned
offsets, k up to the variant's limit, and the "table left all-zero" invariant.  Runs without a GPU; the same source
is what the sm_100a kernels inline."""
import os
import shutil
import subprocess

import pytest

import _oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [1, 2])
def test_bitpar_cores_on_host(tmp_path, seed):
    if not shutil.which("g++"):
        pytest.skip("no g++ on this box")
    orc.build()
    exe = str(tmp_path / "core_host")
    odir = os.path.join(ROOT, "oracle")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-w", os.path.join(ROOT, "tests", "cpp", "core_host.cpp"), "-L", odir,
                           "-lta_oracle", "-Wl,-rpath," + odir, "-o", exe])
    r = subprocess.run([exe, "15000", str(seed)], capture_output=True, text=True)
    assert r.returncode == 0 and " bad 0" in r.stdout, r.stdout[-2000:] + r.stderr[-1000:]
