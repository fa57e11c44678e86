"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol
that include/flof_b200.h declares (no compute calls: there is no GPU here), and it refuses to
create a context without a device instead of falling back to the CPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "flof_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(flof_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from ofblend_b200 import capi
    lib = capi.load_library()
    syms = declared_symbols()
    assert len(syms) >= 50
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_header_cites_reference_for_each_entry_point():
    txt = open(HEADER).read()
    assert txt.count("ref:") >= 40


def test_no_cpu_fallback_without_device():
    from ofblend_b200 import capi
    lib = capi.load_library()
    lib.flof_device_count.restype = ctypes.c_int
    if lib.flof_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.FlofError):
        capi.Context(0)


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/ (checker only)."""
    pkg = os.path.join(ROOT, "ofblend_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "flof_oracle" not in src and "libofref" not in src, f
