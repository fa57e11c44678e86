"""Golden vectors generated from the reference itself by tests/golden/make_golden.py.

CPU part (`not gpu`): the C oracle reproduces every stored reference output.
GPU part: the CUDA path reproduces them through the C ABI (bit-exact except the CG result,
which is held to the 1e-4 relative-L2 bar and must stop at the stored iteration)."""
import os

import numpy as np
import pytest

from conftest import rel_l2, sdf_pair
from oracle import port

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
D = (14, 12, 13, 16)
SH = (D[3], D[2], D[1], D[0])


def rnd(shape, seed, scale=1.0):
    return (np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32)


def eq(a, b):
    assert a.shape == b.shape and np.array_equal(a, b), "max abs diff %g" % np.abs(a.astype(np.float64) - b).max()


def run_small(api, cg_exact):
    g = np.load(os.path.join(G, "small_ops.npz"))
    i0, i1 = sdf_pair(D)
    vel = rnd(SH + (4,), 3, 2.5)
    eq(api.interpolate_grid4d(i0, (7, 6, 6, 8), (1.4, 1.2, 1.3, 3.0), 0.8), g["interp_real"])
    down = api.interpol_grid_templ(vel, (7, 6, 6, 8))
    eq(down, g["interp_vec_down"])
    eq(api.interpol_grid_templ(down, D), g["interp_vec_up"])
    eq(api.advect4d(vel, i0, 0.7), g["advect_real"])
    eq(api.advect4d(vel, rnd(SH + (4,), 4), 0.7), g["advect_vec4"])
    eq(api.advect_cfl4d(1.5, vel, i0, 1.0), g["advect_cfl"])
    eq(api.gaussian_blur4d(rnd(SH + (4,), 9), 2.0), g["blur_s2"])
    eq(api.gaussian_blur4d(rnd(SH + (4,), 9), 1.125), g["blur_s1"])
    v0 = np.zeros(SH + (4,), np.float32)
    of, rhs, it = api.optical_flow4d(v0, i0, i1, 1e-3, 1e-4, 0., 1e-2, -1., want_rhs=True, want_iters=True)
    eq(rhs, g["of_rhs"])
    assert it == int(g["of_iters"][0])
    if cg_exact:
        eq(of, g["of_vel"])
    else:
        assert rel_l2(of, g["of_vel"]) <= 1e-4
    velp = rnd(SH + (4,), 11, 0.7)
    pd, pm = api.project_cells(velp, i0, i1, 4., 40)
    eq(pm, g["proj_marker"])
    eq(pd, g["proj_dst"])
    eq(api.cv_expol_blur4d(pd, pm, 5), g["expol5"])
    cd, cv = api.corr_vels_of4d(np.zeros_like(velp), rnd(SH + (4,), 12, 0.5), i0, i1, 4., 4., 0.1, 40)
    eq(cd, g["corr_dst"])
    eq(cv, g["corr_vel"])
    assert np.allclose([api.calc_ls_diff4d(i0, i1, 0.005, 0), api.calc_ls_diff4d(i0, i1, 0.005, 2)], g["lsdiff"],
                       rtol=1e-6)
    phi = api.set_bound4d(i0 / np.float32(-0.005), 0.1, 1)
    for inside in (0, 1):
        p, m = api.extrap4d_ls_simple(phi, 6, bool(inside), want_marker=True)
        assert m.dtype == np.int32 and np.array_equal(m, g["extrap_marker_%d" % inside])
        eq(p, g["extrap_phi_%d" % inside])
    eq(api.extrapolate_vec4_simple(rnd(SH + (4,), 13), phi, 5), g["extrap_vec4"])
    eq(api.repeat_frame4d(rnd(SH, 14), 4.3, 3.0, 0), g["repeat"])
    eq(api.set_bound_neumann4d(vel, 1), g["neumann_w1"])
    eq(api.set_bound4d(i0, 0.1, 3), g["setbound_w3"])
    eq(api.simple_blur_special(rnd((13, 12, 14), 15), 2, -999., 1), g["blur_special"])


def test_oracle_reproduces_small_goldens():
    run_small(port, cg_exact=True)


@pytest.mark.gpu
def test_gpu_reproduces_small_goldens():
    from ofblend_b200 import capi
    api = capi.HostAPI()
    run_small(api, cg_exact=False)
    api.ctx.close()


def test_grid4dop_reference_test_values():
    """tools/tests/test_0032_grid4dop.py: analytic expectations 1.1/1.2/2.9 on a 10x20x30x12 grid."""
    sh = (12, 30, 20, 10)
    one = np.full(sh, 1.0, np.float32)
    a = port.grid_op4d("addConst", np.zeros(sh, np.float32), None, 1.1)
    assert abs(a - np.float32(1.1)).max() < 1e-7
    b = port.grid_op4d("multConst", one, None, 1.2)
    c = port.grid_op4d("addScaled", a, b, 1.5)            # 1.1 + 1.5*1.2 = 2.9
    assert abs(c - 2.9).max() < 5e-7
    v = port.grid_op4d("add", np.full(sh + (4,), 1.2, np.float32), np.full(sh + (4,), 0.5, np.float32))
    assert abs(v - 1.7).max() < 5e-7


def _mode1_case(name):
    fn = os.path.join(G, name)
    if not os.path.isfile(fn):
        pytest.skip(name + " not generated")
    g = np.load(fn)
    if "vel_noproj_sub" not in g.files:
        pytest.skip(name + " is from an older make_golden.py")
    return g


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mode1_32x48.npz", "mode1_64x64.npz"])
def test_gpu_mode1_matches_reference_run(name):
    """Synthetic two-drop pair at BASELINE.json's sizes: CG iteration counts and the error trace of
    the reference run must be reproduced, the deformation on the stored lattice within 1e-4 rel-L2
    and the applied SDF within 1e-3 cells."""
    g = _mode1_case(name)
    from ofblend_b200 import capi, synth
    api = capi.HostAPI()
    dims = tuple(int(x) for x in g["dims"])
    i0 = synth.post_process(synth.two_drop_phi(dims, 0), api)
    i1 = synth.post_process(synth.two_drop_phi(dims, 1), api)
    assert abs(i0.astype(np.float64).sum() - float(g["i0_sum"])) <= 1e-9 * abs(float(g["i0_sum"])) + 1e-12
    v0 = np.zeros(i0.shape + (4,), np.float32)
    vel, iters, errs = api.optical_flow_multiscale4d(v0, i0, i1, want_trace=True, **synth.MODE1_PARAMS)
    # every CG solve stops at the reference's iteration; the error trace of all OF steps is reproduced
    assert iters == [int(x) for x in g["cg_iters"]], (iters, g["cg_iters"])
    assert np.allclose(errs[:-1], g["errs"][:-1], rtol=2e-5), (errs, g["errs"])
    s = int(g["stride"])
    sub = (slice(None, None, s),) * 4
    # (1) deformation before the final SDF projection
    pnp = dict(synth.MODE1_PARAMS)
    pnp["doFinalProject"] = False
    vel_np = api.optical_flow_multiscale4d(v0, i0, i1, **pnp)
    assert np.array_equal(vel_np[sub], g["vel_noproj_sub"]), rel_l2(vel_np[sub], g["vel_noproj_sub"])
    assert np.array_equal(api.advect4d(vel_np, i0)[sub], g["adv_noproj_sub"])
    # (2) the product's real output, WITH the projection (corrVelsOf4d amplifies round-off level input differences by
    # ~4e4): the CG dot products are summed in the reference's sequential order, so the field -- and the
    # SDF it is applied to (mode 2 on the product's own mode-1 output) -- equal the reference run bit for bit.
    # North-star bars: 1e-4 relative L2 / 1e-3 cells; measured 0 / 0.
    assert np.array_equal(vel[sub], g["vel_sub"]), rel_l2(vel[sub], g["vel_sub"])
    assert np.array_equal(api.advect4d(vel, i0)[sub], g["adv_sub"])
    assert abs(errs[-1] - float(g["errs"][-1])) <= 2e-5 * float(g["errs"][-1])
    api.ctx.close()
