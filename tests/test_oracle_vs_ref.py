"""Pins the C restatement (oracle/flof_oracle.c) against the reference itself
(oracle/_ref/libofref.so = thunil/ofblend compiled unmodified).  CPU only."""
import os

import numpy as np
import pytest

from conftest import rel_l2, sdf_pair
from oracle import port, ref

pytestmark = pytest.mark.ref

D = (14, 12, 13, 16)  # nx, ny, nz, nt (SURVEY.md §8c probe size)
SH = (D[3], D[2], D[1], D[0])


def rnd(shape, seed, scale=1.0):
    return (np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32)


def eq(a, b):
    assert a.shape == b.shape
    assert np.array_equal(a, b), "max abs diff %g" % np.abs(a.astype(np.float64) - b).max()


@pytest.mark.parametrize("elem", [1, 4])
def test_interpolate_grid4d(elem):
    src = rnd(SH + ((4,) if elem == 4 else ()), 1)
    for tdims, off, sc in [((20, 9, 13, 11), 0., 1.), ((7, 6, 6, 8), (1.4, 1.2, 1.3, 3.0), 0.8),
                           ((28, 24, 26, 32), 0., (1., 1., 1., 0.9))]:
        eq(port.interpolate_grid4d(src, tdims, off, sc), ref.interpolate_grid4d(src, tdims, off, sc))


@pytest.mark.parametrize("elem", [1, 4])
def test_interpol_grid_templ(elem):
    src = rnd(SH + ((4,) if elem == 4 else ()), 2)
    down = (D[0] // 2, D[1] // 2, D[2] // 2, D[3] // 2)
    a = port.interpol_grid_templ(src, down)
    eq(a, ref.interpol_grid_templ(src, down))
    eq(port.interpol_grid_templ(a, D), ref.interpol_grid_templ(a, D))


@pytest.mark.parametrize("elem", [1, 4])
def test_advect4d(elem):
    vel = rnd(SH + (4,), 3, 2.5)
    g = rnd(SH + ((4,) if elem == 4 else ()), 4)
    eq(port.advect4d(vel, g, 0.7), ref.advect4d(vel, g, 0.7))
    eq(port.advect_cfl4d(1.5, vel, g, 1.0), ref.advect_cfl4d(1.5, vel, g, 1.0))
    eq(port.advect_cfl4d(999., vel, g, 0.5), ref.advect_cfl4d(999., vel, g, 0.5))


def test_set_bound_and_minmax():
    a = rnd(SH, 5)
    v = rnd(SH + (4,), 6)
    for w in (0, 1, 3):
        eq(port.set_bound4d(a, 0.1, w), ref.set_bound4d(a, 0.1, w))
        eq(port.set_bound4d(v, (1, 2, 3, 4), w), ref.set_bound4d(v, (1, 2, 3, 4), w))
        eq(port.set_bound_neumann4d(a, w), ref.set_bound_neumann4d(a, w))
        eq(port.set_bound_neumann4d(v, w), ref.set_bound_neumann4d(v, w))
    assert port.min_max4d(a) == ref.min_max4d(a)
    assert port.min_max4d(v) == ref.min_max4d(v)


@pytest.mark.parametrize("elem", [1, 4])
def test_grid_ops(elem):
    sh = SH + ((4,) if elem == 4 else ())
    a, b = rnd(sh, 7), rnd(sh, 8)
    for op in ("add", "sub", "mult"):
        eq(port.grid_op4d(op, a, b), ref.grid_op4d(op, a, b))
    f = (0.3, -1.5, 2.0, 0.25) if elem == 4 else 0.3
    eq(port.grid_op4d("addScaled", a, b, f), ref.grid_op4d("addScaled", a, b, f))
    eq(port.grid_op4d("multConst", a, None, f), ref.grid_op4d("multConst", a, None, f))
    eq(port.grid_op4d("addConst", a, None, f), ref.grid_op4d("addConst", a, None, f))
    eq(port.grid_op4d("clamp", a, None, (-0.5, 0.25)), ref.grid_op4d("clamp", a, None, (-0.5, 0.25)))


@pytest.mark.parametrize("sigma", [0.5, 1.0, 1.125, 2.0])
def test_gaussian_blur(sigma):
    v = rnd(SH + (4,), 9)
    eq(port.gaussian_blur4d(v, sigma), ref.gaussian_blur4d(v, sigma))


def test_optical_flow4d_bitexact():
    i0, i1 = sdf_pair(D)
    v0 = np.zeros(SH + (4,), np.float32)
    a, rhs_a, it = port.optical_flow4d(v0, i0, i1, 1e-3, 1e-4, 0., 1e-2, -1., want_rhs=True, want_iters=True)
    b, rhs_b = ref.optical_flow4d(v0, i0, i1, 1e-3, 1e-4, 0., 1e-2, -1., want_rhs=True)
    eq(rhs_a, rhs_b)
    eq(a, b)
    assert it > 3
    # non-zero incoming velocity exercises the smoothness / Tikhonov terms of the rhs
    v1 = rnd(SH + (4,), 10, 0.3)
    eq(port.optical_flow4d(v1, i0, i1, 1e-3, 1e-4, 4., 1e-2, 0.1), ref.optical_flow4d(v1, i0, i1, 1e-3, 1e-4, 4., 1e-2, 0.1))


def test_project_and_expol():
    i0, i1 = sdf_pair(D)
    vel = rnd(SH + (4,), 11, 0.7)
    da, ma = port.project_cells(vel, i0, i1, 4., 40)
    db, mb = ref.project_cells(vel, i0, i1, 4., 40)
    eq(ma, mb)
    eq(da, db)
    assert 0 < ma.sum() < ma.size
    eq(port.cv_expol_blur4d(da, ma, 5), ref.cv_expol_blur4d(db, mb, 5))


def test_corr_vels_of4d():
    i0, i1 = sdf_pair(D)
    vel = rnd(SH + (4,), 12, 0.5)
    z = np.zeros_like(vel)
    a = port.corr_vels_of4d(z, vel, i0, i1, 4., 4., 0.1, 40)
    b = ref.corr_vels_of4d(z, vel, i0, i1, 4., 4., 0.1, 40)
    eq(a[0], b[0])
    eq(a[1], b[1])


def test_calc_ls_diff():
    i0, i1 = sdf_pair(D)
    for bnd in (0, 2):
        ra, oa = port.calc_ls_diff4d(i0, i1, 0.005, bnd, want_out=True)
        rb, ob = ref.calc_ls_diff4d(i0, i1, 0.005, bnd, want_out=True)
        assert ra == rb
        eq(oa, ob)


def test_extrapolation_bitexact_marker():
    i0, _ = sdf_pair(D)
    phi = ref.set_bound4d(i0 / np.float32(-0.005), 0.1, 1)
    for inside in (False, True):
        pa, ma = port.extrap4d_ls_simple(phi, 6, inside, want_marker=True)
        pb, mb = ref.extrap4d_ls_simple(phi, 6, inside, want_marker=True)
        assert ma.dtype == np.int32 and np.array_equal(ma, mb)
        eq(pa, pb)
        eq(pa, ref.extrap4d_ls_simple(phi, 6, inside))
    vel = rnd(SH + (4,), 13)
    eq(port.extrapolate_vec4_simple(vel, phi, 5), ref.extrapolate_vec4_simple(vel, phi, 5))


def test_repeat_frame():
    a = rnd(SH, 14)
    eq(port.repeat_frame4d(a, 4.3, 3.0, 0), ref.repeat_frame4d(a, 4.3, 3.0, 0))
    eq(port.repeat_frame4d(a, 5.0, 2.0, 1), ref.repeat_frame4d(a, 5.0, 2.0, 1))


def test_3d_output_ops():
    a = rnd((13, 12, 14), 15)
    b = rnd((13, 12, 14), 16)
    eq(port.simple_blur_special(a, 1, 0., 2), ref.simple_blur_special(a, 1, 0., 2))
    eq(port.simple_blur_special(a, 2, -999., 1), ref.simple_blur_special(a, 2, -999., 1))
    eq(port.grid3_set_bound(a, 0.5, 2), ref.grid3_set_bound(a, 0.5, 2))
    eq(port.levelset_join(a, b), ref.levelset_join(a, b))


def test_multiscale_small():
    """Full V-cycle (2 levels, 3 steps, final projection) on a 24^3x20 pair: bit-exact."""
    d = (24, 24, 24, 20)
    i0, i1 = sdf_pair(d, seed=3)
    v0 = np.zeros((d[3], d[2], d[1], d[0], 4), np.float32)
    kw = dict(wSmooth=1e-3, wEnergy=1e-4, postVelBlur=4., cgAccuracy=1e-2, cfl=999., resetBndWidth=0.1,
              multiStep=3, minGridSize=20, doFinalProject=True)
    a, iters, errs = port.optical_flow_multiscale4d(v0, i0, i1, want_trace=True, **kw)
    b = ref.optical_flow_multiscale4d(v0, i0, i1, **kw)
    assert len(iters) >= 4 and np.abs(b).max() > 0.05
    eq(a, b)


def test_load_place_and_lookup(tmp_path):
    """Mode-3 loaders: reference reads .uni files it wrote itself, the port gets the arrays."""
    sx = 16
    nfiles = 18
    slices = rnd((nfiles, sx, sx, sx), 17)
    pat = os.path.join(str(tmp_path), "sl_%04d.uni")
    for i in range(nfiles):
        ref.grid3_save(slices[i], pat % i)
    d = (20, 20, 20, 30)
    phi0 = np.zeros((d[3], d[2], d[1], d[0]), np.float32)
    off = (2., 2., 2., 6.)
    sc = (0.8, 0.8, 0.8, 0.8)
    kw = dict(fileIdxStart=0, fileIdxEnd=nfiles, spread=1., overrideSize=(20, 20, 20, 30), rescaleSdfValues=True,
              sdfIsoOff=-0.5, repeatStartFrame=0.1)
    a = port.load_place_grid4d(slices, phi0, off, sc, **kw)
    b = ref.load_place_grid4d(pat, phi0, off, sc, **kw)
    eq(a, b)
    assert np.abs(b).max() > 0
    # partial reload of the tail (streaming window)
    a2 = port.load_place_grid4d(slices, a, off, sc, overrideGoodRegion=22, overrideTimeOff=-3., **kw)
    b2 = ref.load_place_grid4d(pat, b, off, sc, overrideGoodRegion=22, overrideTimeOff=-3., **kw)
    eq(a2, b2)
    eq(port.shift_forw_grid4d(a, 26), ref.shift_forw_grid4d(b, 26))

    # per-frame lookup with on-the-fly velocity up-sampling (K19)
    dd = (8, 8, 8, 12)
    defo = rnd((dd[3], dd[2], dd[1], dd[0], 4), 18, 0.4)
    fn = os.path.join(str(tmp_path), "defo_vel.uni")
    ref.grid4d_save(defo, fn)
    big = (44, 44, 44, 30)
    phi = rnd((big[3], big[2], big[1], big[0]), 19)
    times = [6.5, 7.5, 8.5, 12.5]
    fac = tuple(big[i] / dd[i] for i in range(4))
    outs = ref.load_advect_time_slices_opt(fn, big[:3], phi, times, 0.5, 1., 0., 1., fac, big, 0., 4, 1.)
    for f, tm in enumerate(times):
        o = port.load_advect_time_slice(defo, big[:3], phi, tm, 0.5, 1., 0., 1., fac, big, 0., 4, 1.)
        eq(o, outs[f])
    assert np.abs(outs).max() > 0


@pytest.mark.parametrize("case", ["two", "two_aligned", "three", "two_eof"])
def test_defo_volumes_two_and_three(tmp_path, case):
    """`thirdload`: loadAdvectTimeSlice_OptInit(useDefoVols=True) + _OptAdd + _OptRun (ref optflow4d.cpp:1822-1863 window
    refresh incl. its start-up quirk, :2015-2089 compositions) -- oracle bit-exact against the reference, frame by frame.
    The time list exercises all three refresh branches: first call at t = 0 (only the last window slice gets filled),
    same t, t + 1, and a jump."""
    dd = (8, 7, 8, 20)                                   # dimT 20 -> window of int(20 * 0.2) = 4 slices
    vols = [rnd((dd[3], dd[2], dd[1], dd[0], 4), 30 + q, 0.5) for q in range(3 if case == "three" else 2)]
    fns = []
    for q, v in enumerate(vols):
        fns.append(os.path.join(str(tmp_path), "defo_%s_%d.uni" % (case, q)))
        ref.grid4d_save(v, fns[-1])
    big = (30, 28, 30, 30)
    phi = rnd((big[3], big[2], big[1], big[0]), 41)
    fac = tuple(big[i] / dd[i] for i in range(4))
    times = [0.9, 1.2, 2.4, 3.6, 3.7, 14.5, 16., 29.]
    if case == "two_eof":   # t = 16 (all), 17, 18, 19 one slice at a time: the re-used file handle runs past the last slice
        times = [24.55, 26.2, 27.85, 29.05]
    kw = dict(doAligned=(case == "two_aligned"), partialLoadFac=0.1, overrideSize=big, overrideTimeOff=0.5, bordSkip=3,
              defoAniFac=0.75)
    a = port.load_advect_defovols(vols, big[:3], phi, times, 0.6, 0.3, 0.2, 1., 0., 1., fac, **kw)
    b = ref.load_advect_defovols(fns, big[:3], phi, times, 0.6, 0.3, 0.2, 1., 0., 1., fac, **kw)
    eq(a, b)
    assert np.abs(b).max() > 0
    assert not np.array_equal(b[0], b[2])


def test_unoptimised_load_advect_time_slice(tmp_path):
    """The slower twin loadAdvectTimeSlice (ref optflow4d.cpp:1671-1760) with its debugVel / debugVelT outputs.  Its
    look-up kernel is KERNEL(fourd, bnd = 1) on a ONE-slice 4D grid: the generated loop runs it as a 3D kernel with t = 0
    over the interior cells, so dst is written everywhere but on the outer shell."""
    dd = (8, 9, 7, 12)
    defo = rnd((dd[3], dd[2], dd[1], dd[0], 4), 51, 0.6)
    fn = os.path.join(str(tmp_path), "defo_dbg.uni")
    ref.grid4d_save(defo, fn)
    big = (21, 19, 20, 30)
    phi = rnd((big[3], big[2], big[1], big[0]), 52)
    fac = tuple(big[i] / dd[i] for i in range(4))
    for tm, zero in ((6.5, False), (11.25, False), (29., False), (8., True)):
        a = port.load_advect_time_slice_unopt(defo, big[:3], phi, tm, 0.7, 1., 0., 1., fac, big, -0.5, 0.8, zero)
        b = ref.load_advect_time_slice_unopt(fn, big[:3], phi, tm, 0.7, 1., 0., 1., fac, big, -0.5, 0.8, zero)
        for x, y in zip(a, b):
            eq(x, y)
        assert np.abs(b[0]).max() > 0
    assert np.abs(b[1]).max() == 0 and np.abs(a[2]).max() == 0     # zeroVel
