"""CPU check of the scheme behind the CG's sequential-order dot products (ofblend_b200/csrc/flof_seqsum_core.h).

The reference sums `d += a[i]*b[i]` one product after the other (optflow4d.cpp:234-241).  The CUDA path evaluates
that same sum in parallel -- per-binade rounding functions, tie parity, raw adds around binade crossings -- and must
return the identical fp64 bits.  tests/seqsum_model.cpp organises the work like the kernels do, on the host, with the
arithmetic header the kernels include; here it is compared with the plain loop on inputs chosen to hit ties, binade
crossings in both directions, zero stretches, huge dynamic range and non-finite values.  The approximate prefix that
drives the classification is also pushed to the edge of its margin (`perturb`) -- the bits must not move."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("seqsum") / "libseqsum_model.so")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so,
                           os.path.join(HERE, "seqsum_model.cpp")])
    L = C.CDLL(so)
    fp = C.POINTER(C.c_float)
    L.seqsum_reference.restype = C.c_double
    L.seqsum_reference.argtypes = [fp, fp, C.c_longlong]
    L.seqsum_model.restype = C.c_double
    L.seqsum_model.argtypes = [fp, fp, C.c_longlong, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_longlong)]
    return L


def run(L, a, b, leaf=4096, thread=16, perturb=0.0, s_in=0.0):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    fp = C.POINTER(C.c_float)
    st = (C.c_longlong * 4)()
    ref = L.seqsum_reference(a.ctypes.data_as(fp), b.ctypes.data_as(fp), a.size) if s_in == 0.0 else None
    got = L.seqsum_model(a.ctypes.data_as(fp), b.ctypes.data_as(fp), a.size, leaf, thread, perturb, s_in, st)
    return ref, got, list(st)


def same_bits(x, y):
    return np.float64(x).tobytes() == np.float64(y).tobytes() or (np.isnan(x) and np.isnan(y))


def cases():
    rng = np.random.default_rng(7)
    n = 300000
    out = {}
    out["normal"] = (rng.standard_normal(n), rng.standard_normal(n))
    out["wide_range"] = (rng.standard_normal(n) * 10.0 ** rng.uniform(-9, 2, n), rng.standard_normal(n))
    r = rng.standard_normal(n) * 10.0 ** rng.uniform(-6, 0, n)
    out["positive_terms"] = (r, r)                                   # sigma = sum r*r*pc: monotone running sum
    z = rng.standard_normal(n)
    z[: n // 3] = 0.0
    z[n // 2: n // 2 + 50000] = 0.0
    out["zero_stretches"] = (z, rng.standard_normal(n))
    # running sum hovering around a power of two: crossings in both directions
    a = np.concatenate([[1.0], rng.standard_normal(n - 1) * 1e-6])
    out["hover_at_one"] = (a, np.ones(n))
    # exact ties: products with a single low bit, 2^-30 .. 2^-53 below the running sum
    t = 2.0 ** -rng.integers(28, 56, n).astype(np.float64)
    out["ties"] = (np.concatenate([[1.5], t[1:]]), np.ones(n))
    out["ties_signed"] = (np.concatenate([[1.5], t[1:] * rng.choice([-1.0, 1.0], n - 1)]), np.ones(n))
    # growing magnitudes: many upward crossings
    out["growing"] = (np.exp(np.linspace(-40, 5, n)), rng.uniform(0.5, 1.5, n))
    # cancellation: large terms that cancel, then small ones
    c = rng.standard_normal(n) * 1e-3
    c[1000], c[2000] = 1e6, -1e6
    out["cancel"] = (c, np.ones(n))
    out["tiny"] = (rng.standard_normal(1000) * 1e-30, rng.standard_normal(1000) * 1e-30)
    return out


@pytest.mark.parametrize("name", list(cases().keys()))
def test_model_matches_sequential_loop(lib, name):
    a, b = cases()[name]
    for leaf, thread in ((4096, 16), (1024, 4), (256, 0)):
        for perturb in (0.0, 1.0):
            ref, got, st = run(lib, a, b, leaf, thread, perturb)
            assert same_bits(ref, got), (name, leaf, thread, perturb, ref, got, st)
            assert st[3] == 0, (name, st)


def test_most_leaves_are_clean(lib):
    """The scheme is only fast if binade crossings are rare: a CG-like dot product must be almost all clean leaves."""
    rng = np.random.default_rng(3)
    n = 1 << 21
    r = (rng.standard_normal(n) * 10.0 ** rng.uniform(-4, 0, n)).astype(np.float32)
    ref, got, st = run(lib, r, r, 4096, 16)
    assert same_bits(ref, got)
    assert st[0] <= 40 and st[1] <= 400, st     # dirty leaves, raw products (of 512 leaves / 2M products)


def test_chained_ranks(lib):
    """Multi-GPU form: each rank continues from the exact running sum of the rank before it."""
    rng = np.random.default_rng(11)
    n = 200000
    a = (rng.standard_normal(n) * 10.0 ** rng.uniform(-5, 1, n)).astype(np.float32)
    b = rng.standard_normal(n).astype(np.float32)
    ref, _, _ = run(lib, a, b)
    s = None
    for part in range(4):
        sl = slice(part * n // 4, (part + 1) * n // 4)
        s = run(lib, a[sl], b[sl])[1] if part == 0 else run(lib, a[sl], b[sl], 4096, 16, 0.0, s)[1]
    assert same_bits(ref, s)


def test_non_finite(lib):
    a = np.ones(5000, np.float32)
    a[1234] = np.inf
    ref, got, _ = run(lib, a, np.ones(5000, np.float32))
    assert np.isinf(ref) and np.isinf(got)
    a[4000] = -np.inf
    ref, got, _ = run(lib, a, np.ones(5000, np.float32))
    assert np.isnan(ref) and np.isnan(got)
