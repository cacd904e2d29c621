"""The thread-per-pair u16 routine of the general-cost kernel (triple_accel_b200/csrc/lev_diag16_core.cuh is host/device
code) compiled for the host -- DPX min/add-min, PRMT and funnel shifts emulated -- and pinned to the scalar oracle:
all ten cost models (weighted, affine, transpositions), every register count that holds the band (2, 3, 4, 6, 8 packed
registers per anti-diagonal), the affine code path also for start_gap = 0, alphabets of 2..256 symbols, ragged lengths,
misaligned offsets, k from 0 to 40.  Runs without a GPU; the same source is what the sm_100a kernel inlines."""
import os
import shutil
import subprocess

import pytest

import _oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [1, 2])
def test_diag16_core_on_host(tmp_path, seed):
    if not shutil.which("g++"):
        pytest.skip("no g++ on this box")
    orc.build()
    exe = str(tmp_path / "diag16_host")
    odir = os.path.join(ROOT, "oracle")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-w", os.path.join(ROOT, "tests", "cpp", "diag16_host.cpp"), "-L", odir,
                           "-lta_oracle", "-Wl,-rpath," + odir, "-o", exe])
    r = subprocess.run([exe, "25000", str(seed)], capture_output=True, text=True)
    assert r.returncode == 0 and " bad 0" in r.stdout, r.stdout[-2000:] + r.stderr[-1000:]
