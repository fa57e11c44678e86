"""Parity of the 3D instantiations (opticalFlowMultiscale3d, corrVelsOf3d, advectSemiLagrangeCfl,
calcLsDiff3d) through the C ABI.  Checkers: the reference's outputs committed as tests/golden/dim3_ops.npz
(tests/golden/make_golden3d.py), the plain-C restatement oracle/flof_oracle3.c on further inputs, and, where
oracle/_ref/libofref.so travelled to the box, the compiled reference run live.  Bar: bit-exact (all fp32 arithmetic follows the
reference's order; the CG runs through the 4D kernels with the reference's sequential-order dot products)."""
import os

import numpy as np
import pytest

from conftest import DIM3_CASES, sdf_pair3
from oracle import ref

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dim3_ops.npz"))
D = (20, 18, 16)
SH = (D[2], D[1], D[0])


@pytest.fixture(scope="module")
def gpu():
    from ofblend_b200 import capi
    api = capi.HostAPI()
    yield api
    api.ctx.close()


def rnd(shape, seed, scale=1.0):
    return (np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32)


def eq(a, b):
    assert a.shape == b.shape
    assert np.array_equal(a, b), "max abs diff %g" % np.abs(a.astype(np.float64) - b).max()


def test_calc_ls_diff3d(gpu):
    i0, i1 = sdf_pair3(D)
    assert np.float32(gpu.calc_ls_diff3d(i0, i1, 0.005 * 200, 0)) == GOLD["lsdiff"][0]
    r, out = gpu.calc_ls_diff3d(i0, i1, 0.005 * 200, 2, want_out=True)
    assert np.float32(r) == GOLD["lsdiff"][1]
    inner = (slice(2, -2),) * 3                        # `out` is only written inside the bound (ref :895-913)
    eq(out[inner], GOLD["lsdiff_out"][inner])


def test_advect_semi_lagrange_cfl3d(gpu):
    i0, _ = sdf_pair3(D)
    vel = rnd(SH + (3,), 3, 2.0)
    eq(gpu.advect_semi_lagrange_cfl3d(999., vel, i0), GOLD["adv_real"])
    eq(gpu.advect_semi_lagrange_cfl3d(999., vel, rnd(SH + (3,), 4)), GOLD["adv_vec3"])
    eq(gpu.advect_semi_lagrange_cfl3d(1.0, vel, i0), GOLD["adv_real_cfl1"])          # CFL sub-steps
    eq(gpu.advect_semi_lagrange_cfl3d(1.5, vel, i0, 0.37), GOLD["adv_real_fac"])     # velFactor


def test_corr_vels_of3d(gpu):
    i0, i1 = sdf_pair3(D)
    d, v = gpu.corr_vels_of3d(np.zeros(SH + (3,), np.float32), rnd(SH + (3,), 12, 0.5), i0, i1, 4., 2., 0.1, 40)
    eq(d, GOLD["corr_dst"])
    eq(v, GOLD["corr_vel"])


@pytest.mark.parametrize("name", sorted(DIM3_CASES))
def test_optical_flow_multiscale3d(gpu, name):
    dims, params = DIM3_CASES[name]
    a, b = sdf_pair3(dims)
    v0 = np.zeros(a.shape + (3,), np.float32)
    v, iters, errs = gpu.optical_flow_multiscale3d(v0, a, b, want_trace=True, **params)
    assert iters == list(GOLD["ms_%s_iters" % name])
    np.testing.assert_allclose(errs, GOLD["ms_%s_errs" % name], rtol=2e-6)
    eq(v, GOLD["ms_%s_vel" % name])
    eq(gpu.advect_semi_lagrange_cfl3d(999., v, a), GOLD["ms_%s_adv" % name])          # the applied SDF


@pytest.mark.ref
def test_dim3_against_the_live_reference(gpu):
    """Further inputs (odd sizes, other seeds) against the compiled reference on the box's host cores."""
    for dims, seed in (((17, 15, 13), 1), ((22, 12, 26), 2), ((26, 23, 1), 3)):      # the last one is a 2D grid
        sh = (dims[2], dims[1], dims[0])
        i0, i1 = sdf_pair3(dims, seed)
        vel = rnd(sh + (3,), 30 + seed, 1.5)
        eq(gpu.advect_semi_lagrange_cfl3d(0.8, vel, i1), ref.advect_semi_lagrange_cfl3d(0.8, vel, i1))
        assert np.float32(gpu.calc_ls_diff3d(i0, i1, 1.0, 1)) == np.float32(ref.calc_ls_diff3d(i0, i1, 1.0, 1))
        gd, gv = gpu.corr_vels_of3d(np.zeros(sh + (3,), np.float32), vel * np.float32(0.3), i0, i1, 4., 4., 0.1, 40)
        rd, rv = ref.corr_vels_of3d(np.zeros(sh + (3,), np.float32), vel * np.float32(0.3), i0, i1, 4., 4., 0.1, 40)
        eq(gd, rd)
        eq(gv, rv)
        p = dict(wSmooth=1e-2, wEnergy=1e-4, postVelBlur=3., cgAccuracy=1e-3, resetBndWidth=0.1, multiStep=2, minGridSize=10,
                 doFinalProject=True)
        v0 = np.zeros(sh + (3,), np.float32)
        eq(gpu.optical_flow_multiscale3d(v0, i0, i1, **p), ref.optical_flow_multiscale3d(v0, i0, i1, **p))


def test_dim3_against_the_c_restatement(gpu):
    """Further inputs against oracle/flof_oracle3.c (pinned bit for bit to the reference by tests/test_dim3_golden.py);
    needs no compiled reference on the box."""
    from oracle import port
    for dims, seed in (((19, 14, 12), 5), ((24, 21, 1), 6)):
        sh = (dims[2], dims[1], dims[0])
        i0, i1 = sdf_pair3(dims, seed)
        vel = rnd(sh + (3,), 40 + seed, 1.5)
        eq(gpu.advect_semi_lagrange_cfl3d(0.7, vel, i1), port.advect_semi_lagrange_cfl3d(0.7, vel, i1))
        eq(gpu.advect_semi_lagrange_cfl3d(999., vel, vel, 0.5), port.advect_semi_lagrange_cfl3d(999., vel, vel, 0.5))
        assert np.float32(gpu.calc_ls_diff3d(i0, i1, 1.0, 1)) == np.float32(port.calc_ls_diff3d(i0, i1, 1.0, 1))
        z = np.zeros(sh + (3,), np.float32)
        gd, gv = gpu.corr_vels_of3d(z, vel * np.float32(0.3), i0, i1, 4., 4., 0.1, 40)
        pd, pv = port.corr_vels_of3d(z, vel * np.float32(0.3), i0, i1, 4., 4., 0.1, 40)
        eq(gd, pd)
        eq(gv, pv)
        p = dict(wSmooth=5e-3, wEnergy=1e-4, postVelBlur=2., cgAccuracy=1e-3, resetBndWidth=0.1, multiStep=3, minGridSize=10,
                 doFinalProject=True)
        a, it_a, err_a = gpu.optical_flow_multiscale3d(z, i0, i1, want_trace=True, **p)
        b, it_b, err_b = port.optical_flow_multiscale3d(z, i0, i1, want_trace=True, **p)
        assert it_a == it_b
        eq(a, b)
