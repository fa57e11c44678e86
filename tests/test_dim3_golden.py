"""The committed 3D fixture (tests/golden/dim3_ops.npz) against the live compiled reference: the GPU tests of
the 3D instantiations compare with this file, so it must be exactly what the reference computes from the seeded inputs."""
import os

import numpy as np
import pytest

from conftest import DIM3_CASES, sdf_pair3
from oracle import ref

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dim3_ops.npz"))


def rnd(shape, seed, scale=1.0):
    return (np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32)


@pytest.mark.ref
def test_operator_fixtures_are_the_references_outputs():
    D = (20, 18, 16)
    SH = (D[2], D[1], D[0])
    i0, i1 = sdf_pair3(D)
    vel = rnd(SH + (3,), 3, 2.0)
    assert np.float32(ref.calc_ls_diff3d(i0, i1, 0.005 * 200, 0)) == GOLD["lsdiff"][0]
    assert np.array_equal(ref.advect_semi_lagrange_cfl3d(1.0, vel, i0), GOLD["adv_real_cfl1"])
    assert np.array_equal(ref.advect_semi_lagrange_cfl3d(999., vel, rnd(SH + (3,), 4)), GOLD["adv_vec3"])
    d, v = ref.corr_vels_of3d(np.zeros(SH + (3,), np.float32), rnd(SH + (3,), 12, 0.5), i0, i1, 4., 2., 0.1, 40)
    assert np.array_equal(d, GOLD["corr_dst"]) and np.array_equal(v, GOLD["corr_vel"])


@pytest.mark.ref
@pytest.mark.parametrize("name", sorted(DIM3_CASES))
def test_multiscale3d_fixture_is_the_references_output(name):
    dims, params = DIM3_CASES[name]
    a, b = sdf_pair3(dims)
    v = ref.optical_flow_multiscale3d(np.zeros(a.shape + (3,), np.float32), a, b, **params)
    assert np.array_equal(v, GOLD["ms_%s_vel" % name])
    assert np.abs(v).max() > 0.5                                   # a real deformation, not a trivial fixture
    # serial and OpenMP runs of the reference agree bit for bit
    n = ref.set_threads(1)
    try:
        assert np.array_equal(ref.optical_flow_multiscale3d(np.zeros(a.shape + (3,), np.float32), a, b, **params), v)
    finally:
        ref.set_threads(os.cpu_count() or 1)


# ---- the plain-C restatement (oracle/flof_oracle3.c) against the compiled reference and against the fixture ----------
@pytest.mark.ref
@pytest.mark.parametrize("dims,seed", [((20, 18, 16), 0), ((17, 15, 13), 1), ((22, 12, 26), 2), ((26, 23, 1), 3)])
def test_oracle3_operators_match_the_reference_bit_for_bit(dims, seed):
    from oracle import port
    sh = (dims[2], dims[1], dims[0])
    i0, i1 = sdf_pair3(dims, seed)
    vel = rnd(sh + (3,), 3 + seed, 2.0)
    for args in ((999., vel, i0), (1.0, vel, i0), (1.5, vel, i0, 0.37), (0.8, vel, rnd(sh + (3,), 4))):
        assert np.array_equal(port.advect_semi_lagrange_cfl3d(*args), ref.advect_semi_lagrange_cfl3d(*args))
    for bnd in (0, 2):
        a = port.calc_ls_diff3d(i0, i1, 1.0, bnd, want_out=True)
        b = ref.calc_ls_diff3d(i0, i1, 1.0, bnd, want_out=True)
        assert np.float32(a[0]) == np.float32(b[0]) and np.array_equal(a[1], b[1])
    z = np.zeros(sh + (3,), np.float32)
    for blur in (2., 0.5):
        a = port.corr_vels_of3d(z, rnd(sh + (3,), 12, 0.5), i0, i1, 4., blur, 0.1, 40)
        b = ref.corr_vels_of3d(z, rnd(sh + (3,), 12, 0.5), i0, i1, 4., blur, 0.1, 40)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    p = dict(wSmooth=1e-2, wEnergy=1e-4, postVelBlur=3., cgAccuracy=1e-3, resetBndWidth=0.1, multiStep=2, minGridSize=10,
             doFinalProject=True)
    assert np.array_equal(port.optical_flow_multiscale3d(z, i0, i1, **p), ref.optical_flow_multiscale3d(z, i0, i1, **p))


@pytest.mark.parametrize("name", sorted(DIM3_CASES))
def test_oracle3_reproduces_the_committed_fixture(name):
    """runs everywhere (no compiled reference needed): restatement == the reference's outputs committed as goldens"""
    from oracle import port
    dims, params = DIM3_CASES[name]
    a, b = sdf_pair3(dims)
    v, iters, errs = port.optical_flow_multiscale3d(np.zeros(a.shape + (3,), np.float32), a, b, want_trace=True, **params)
    assert np.array_equal(v, GOLD["ms_%s_vel" % name])
    assert iters == list(GOLD["ms_%s_iters" % name])
    np.testing.assert_allclose(errs, GOLD["ms_%s_errs" % name], rtol=2e-6)
    assert np.array_equal(port.advect_semi_lagrange_cfl3d(999., v, a), GOLD["ms_%s_adv" % name])
