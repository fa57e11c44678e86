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


def test_manta_module_loads_and_refuses_to_compute_without_device(tmp_path):
    """The C++ host layer imports on a CPU box (the device context is created on first use), answers pure file
    queries through the same I/O pool the GPU path uses, and fails loudly -- no CPU fallback -- as soon as a grid
    is requested without a device."""
    import subprocess
    import sys
    host = os.path.join(ROOT, "ofblend_b200", "host")
    if not [f for f in os.listdir(host) if f.startswith("manta") and f.endswith(".so")]:
        pytest.skip("host module not built")
    code = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import manta
from ofblend_b200 import uni
fn = os.path.join(sys.argv[3], "a.uni")
uni.write_uni(fn, np.arange(5 * 4 * 3, dtype=np.float32).reshape(5, 4, 3))
v = manta.getUniFileSize(fn)
assert (v.x, v.y, v.z) == (3.0, 4.0, 5.0), (v.x, v.y, v.z)
manta.flushUniWrites()
for name in ("opticalFlowMultiscale4d", "advect4d", "loadAdvectTimeSlice_OptRun", "corrVelsOf4d", "extrap4dLsSimple",
             "loadPlaceGrid4d", "calcObfDiff", "debugVelAvg4d", "grid4dMaxDiffInt", "Grid4Real", "Grid4Vec4", "Solver"):
    assert hasattr(manta, name), name
from ofblend_b200 import capi
lib = capi.load_library()
lib.flof_device_count.restype = int
if lib.flof_device_count() == 0:
    try:
        s = manta.Solver(name="s", gridSize=manta.vec3(4, 4, 4), dim=3)
        s.create(manta.RealGrid)
    except RuntimeError as e:
        assert "no CUDA device" in str(e), e
    else:
        raise SystemExit("a grid was created without a CUDA device")
print("OK")
'''
    p = subprocess.run([sys.executable, "-c", code, host, ROOT, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "OK" in p.stdout, p.stdout[-1500:] + p.stderr[-1500:]
