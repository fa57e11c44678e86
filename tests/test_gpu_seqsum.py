"""GPU: the CG dot products in the reference's sequential summation order (flof_dot_seq, dot_mode 1).

The oracle is the reference's own loop `for (i) d += a[i]*b[i]` (oracle/flof_oracle.c orc_dot_seq, ref optflow4d.cpp:234-241);
the CUDA path must return the identical fp64 BITS and must get there without falling back to its one-thread loop.
With every dot product exact, a whole mode-1 solve -- CG, blurs, projection -- is bit-identical to the reference."""
import os

import numpy as np
import pytest

from conftest import sdf_pair
from oracle import port

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gpu():
    from ofblend_b200 import capi
    api = capi.HostAPI()
    yield api
    api.ctx.close()


def bits(x):
    return np.float64(x).tobytes()


def vec_cases(cells):
    rng = np.random.default_rng(5)
    n = cells * 4
    out = {}
    r = (rng.standard_normal(n) * 10.0 ** rng.uniform(-6, 0, n)).astype(np.float32)
    out["positive_terms"] = (r, r)
    out["normal"] = (rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32))
    z = rng.standard_normal(n).astype(np.float32)
    z[: n // 3] = 0.0
    out["zero_stretches"] = (z, np.abs(rng.standard_normal(n)).astype(np.float32) * z)
    t = (2.0 ** -rng.integers(28, 56, n).astype(np.float64)).astype(np.float32)
    t[0] = 1.5
    out["ties"] = (t, np.ones(n, np.float32))
    out["growing"] = (np.exp(np.linspace(-40, 5, n)).astype(np.float32), rng.uniform(0.5, 1.5, n).astype(np.float32))
    a = (rng.standard_normal(n) * 1e-6).astype(np.float32)
    a[0] = 1.0
    out["hover_at_one"] = (a, np.ones(n, np.float32))
    return out


@pytest.mark.parametrize("cells", [1, 255, 1024, 5000, 70001, 300000])
def test_dot_seq_bits_small(gpu, cells):
    for name, (a, b) in vec_cases(cells).items():
        da = gpu.ctx.to_device(a.reshape(1, 1, 1, cells, 4))
        db = gpu.ctx.to_device(b.reshape(1, 1, 1, cells, 4))
        got, st = gpu.ctx.dot_seq(da, db, 0)
        want = port.dot_seq(a, b, 0)
        assert bits(got) == bits(want), (name, cells, got, want, st)
        da.free()
        db.free()
    st = gpu.ctx.seq_stats()
    assert st["inconsistent"] == 0, st


def test_dot_seq_preconditioned_and_full_size(gpu):
    """kind 1 (tmp = res*precond; dot(tmp, res)) with identity-row markers and kind 0 on a CG-like pair (sum dominated by
    one sign, like srch . A srch), at the 64^4 cell count: exact bits, carried by the parallel scheme (no fallback)."""
    rng = np.random.default_rng(9)
    cells = 64 ** 4
    res = (rng.standard_normal((cells, 4)) * 10.0 ** rng.uniform(-5, -1, (cells, 1))).astype(np.float32)
    grad = rng.standard_normal((cells, 4)).astype(np.float32) * np.float32(0.3)
    grad[:: 97, 0] = np.nan            # identity rows (border cells)
    res[: 64 ** 3] = 0.0               # first t-slice: border, zero residual
    diag = np.float32(0.0081)
    da = gpu.ctx.to_device(res.reshape(1, 1, 1, cells, 4))
    db = gpu.ctx.to_device(grad.reshape(1, 1, 1, cells, 4))
    before = gpu.ctx.seq_stats()
    got, _ = gpu.ctx.dot_seq(da, db, 1, diag)
    want = port.dot_seq(res, grad, 1, diag)
    assert bits(got) == bits(want), (got, want)
    got0, _ = gpu.ctx.dot_seq(da, db, 0)           # kind 0 multiplies the NaN markers in: NaN on both sides (payload free)
    assert np.isnan(got0) and np.isnan(port.dot_seq(res, grad, 0))
    # srch . A srch of an SPD system: mostly positive products, some negative ones
    ap = (res * (1.0 + 0.5 * rng.standard_normal((cells, 4)))).astype(np.float32)
    dc = gpu.ctx.to_device(ap.reshape(1, 1, 1, cells, 4))
    got0, _ = gpu.ctx.dot_seq(da, dc, 0)
    assert bits(got0) == bits(port.dot_seq(res, ap, 0))
    after = gpu.ctx.seq_stats()
    assert after["fallbacks"] == before["fallbacks"] and after["inexact"] == before["inexact"], (before, after)
    # a monotone sum crosses each binade once: only a handful of the 16384 leaves may be dirty
    assert after["dirty_leaves"] - before["dirty_leaves"] <= 400, (before, after)
    for g in (da, db, dc):
        g.free()


def test_dot_seq_heavy_cancellation_falls_back_exactly(gpu):
    """A sum that hovers around zero (random signs) is outside what the CG produces; up to 2^21 cells the resolver's
    fallback is the plain loop -- still the exact bits -- and it is counted."""
    rng = np.random.default_rng(11)
    cells = 1 << 20
    a = rng.standard_normal((cells, 4)).astype(np.float32)
    b = rng.standard_normal((cells, 4)).astype(np.float32)
    da = gpu.ctx.to_device(a.reshape(1, 1, 1, cells, 4))
    db = gpu.ctx.to_device(b.reshape(1, 1, 1, cells, 4))
    got, _ = gpu.ctx.dot_seq(da, db, 0)
    assert bits(got) == bits(port.dot_seq(a, b, 0))
    st = gpu.ctx.seq_stats()
    assert st["inexact"] == 0 and st["inconsistent"] == 0, st
    da.free()
    db.free()


def test_solve_is_bit_identical_to_the_oracle(gpu):
    """opticalFlow4d (assembly + CG + blur + border reset) against the oracle in the REFERENCE's summation order."""
    d = (24, 20, 22, 18)
    i0, i1 = sdf_pair(d, seed=5)
    v0 = np.zeros((d[3], d[2], d[1], d[0], 4), np.float32)
    before = gpu.ctx.seq_stats()
    a, it_a = gpu.optical_flow4d(v0, i0, i1, 1e-3, 1e-4, 4., 1e-2, 0.1, want_iters=True)
    b, it_b = port.optical_flow4d(v0, i0, i1, 1e-3, 1e-4, 4., 1e-2, 0.1, want_iters=True)
    assert it_a == it_b
    assert np.array_equal(a, b), np.abs(a - b).max()
    assert gpu.ctx.seq_stats()["fallbacks"] == before["fallbacks"]


def test_mode1_small_is_bit_identical_to_the_oracle(gpu):
    """Full V-cycle with final projection: identical bits, not a tolerance."""
    d = (24, 24, 24, 20)
    i0, i1 = sdf_pair(d, seed=3)
    v0 = np.zeros((d[3], d[2], d[1], d[0], 4), np.float32)
    kw = dict(wSmooth=1e-3, wEnergy=1e-4, postVelBlur=4., cgAccuracy=1e-2, cfl=999., resetBndWidth=0.1,
              multiStep=3, minGridSize=20, doFinalProject=True)
    a, it_a, err_a = gpu.optical_flow_multiscale4d(v0, i0, i1, want_trace=True, **kw)
    b, it_b, err_b = port.optical_flow_multiscale4d(v0, i0, i1, want_trace=True, **kw)
    assert it_a == it_b
    assert np.array_equal(a, b), np.abs(a - b).max()
    assert np.allclose(err_a, err_b, rtol=1e-6)


@pytest.mark.parametrize("name", ["mode1_32x48.npz", "mode1_64x64.npz"])
def test_mode1_matches_the_reference_run_bit_for_bit(gpu, name):
    """BASELINE.json sizes (32^3 x 48, 64^4), README parameters, WITH the final SDF projection: the deformation the product
    returns equals the reference run's on the stored lattice bit for bit, the applied SDF likewise (north-star bars:
    1e-4 relative L2 / 1e-3 cells -- met with 0)."""
    from ofblend_b200 import synth
    g = np.load(os.path.join(G, name))
    dims = tuple(int(x) for x in g["dims"])
    before = gpu.ctx.seq_stats()
    i0 = synth.post_process(synth.two_drop_phi(dims, 0), gpu)
    i1 = synth.post_process(synth.two_drop_phi(dims, 1), gpu)
    v0 = np.zeros(i0.shape + (4,), np.float32)
    vel, iters, errs = gpu.optical_flow_multiscale4d(v0, i0, i1, want_trace=True, **synth.MODE1_PARAMS)
    assert iters == [int(x) for x in g["cg_iters"]], (iters, g["cg_iters"])
    s = int(g["stride"])
    sub = (slice(None, None, s),) * 4
    assert np.array_equal(vel[sub], g["vel_sub"]), np.abs(vel[sub] - g["vel_sub"]).max()
    # (numpy's pairwise summation is not reproducible across builds to the last bits)
    assert abs(np.linalg.norm(vel.astype(np.float64).ravel()) - float(g["vel_l2"])) <= 1e-12 * float(g["vel_l2"])
    adv = gpu.advect4d(vel, i0)          # mode 2 on the product's own mode-1 output
    assert np.array_equal(adv[sub], g["adv_sub"]), np.abs(adv[sub] - g["adv_sub"]).max() / 0.005
    assert np.allclose(errs, g["errs"], rtol=1e-6), (errs, g["errs"])
    st = gpu.ctx.seq_stats()
    assert st["fallbacks"] == before["fallbacks"] and st["inconsistent"] == 0 and st["inexact"] == 0, (before, st)


def test_dot_seq_fragmented_leaves_are_summed_product_by_product(gpu):
    """A large sum (> 2^21 cells: no plain-loop fallback) whose running value hovers around zero for its first 48 leaves:
    those leaves are kept as plain products (raw leaves) and added one by one by the resolver -- exact bits, no fallback."""
    rng = np.random.default_rng(23)
    cells = (1 << 22) + 4321
    a = np.abs(rng.standard_normal((cells, 4))).astype(np.float32)
    b = np.abs(rng.standard_normal((cells, 4))).astype(np.float32)
    n0 = 48 * 1024
    a[:n0] *= rng.choice(np.array([-1.0, 1.0], np.float32), size=(n0, 4))
    da = gpu.ctx.to_device(a.reshape(1, 1, 1, cells, 4))
    db = gpu.ctx.to_device(b.reshape(1, 1, 1, cells, 4))
    before = gpu.ctx.seq_stats()
    got, _ = gpu.ctx.dot_seq(da, db, 0)
    assert bits(got) == bits(port.dot_seq(a, b, 0))
    st = gpu.ctx.seq_stats()
    assert st["fallbacks"] == before["fallbacks"] and st["inexact"] == 0 and st["inconsistent"] == 0, (before, st)
    assert st["raw_leaves"] > before["raw_leaves"], (before, st)
    da.free()
    db.free()
