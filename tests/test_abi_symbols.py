"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
and fails loudly (no fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from triple_accel_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "triple_accel_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ta_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _ffi.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), "libtriple_accel_b200.so does not export %s" % name
    # and the Python binding declares a signature for each of them
    assert set(names) == set(_ffi.SIGNATURES)


def test_abi_version_and_strerror():
    lib = _ffi.load()
    assert lib.ta_abi_version() == 2
    assert lib.ta_strerror(0) == b"ok"
    assert b"length" in lib.ta_strerror(_ffi.TA_ERR_LEN_MISMATCH)


def test_struct_layouts_match_header():
    assert C.sizeof(_ffi.ta_costs) == 4
    assert C.sizeof(_ffi.ta_match) == 24
    assert _ffi.ta_match.k.offset == 16


def test_costs_validation_matches_reference_asserts():
    lib = _ffi.load()
    ok = lambda *c: lib.ta_costs_valid(_ffi.ta_costs(*c))
    oks = lambda *c: lib.ta_costs_valid_search(_ffi.ta_costs(*c))
    assert ok(1, 1, 0, 0) and ok(1, 1, 0, 1) and ok(2, 3, 0, 0) and ok(1, 1, 2, 0)
    assert not ok(0, 1, 0, 0) and not ok(1, 0, 0, 0)  # src/levenshtein.rs:44-45
    assert not ok(1, 1, 0, 2) and ok(2, 2, 0, 3) and not ok(2, 1, 0, 2)  # :50-51
    assert oks(1, 1, 0, 1) and not oks(2, 2, 0, 3) and oks(2, 2, 1, 3)  # :69


def test_search_default_k():
    lib = _ffi.load()
    assert [lib.ta_search_default_k(n) for n in (0, 1, 2, 3, 32, 33)] == [0, 1, 1, 2, 16, 17]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from triple_accel_b200 import Engine, TripleAccelError
    with pytest.raises(TripleAccelError):
        Engine(0)


def test_python_editcosts_asserts():
    from triple_accel_b200 import EditCosts
    EditCosts(1, 1, 0, 1)
    with pytest.raises(AssertionError):
        EditCosts(0, 1, 0)
    with pytest.raises(AssertionError):
        EditCosts(1, 1, 0, 2)
    with pytest.raises(AssertionError):
        EditCosts(2, 2, 0, 3).check_search()
