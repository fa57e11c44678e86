"""Parity of the CUDA path (through the C ABI, ofblend_b200/capi.py) against the CPU oracle
(oracle/flof_oracle.c, pinned bit-for-bit to the reference by tests/test_oracle_vs_ref.py).

Bars (BASELINE.json north_star): integer/flag grids bit-exact; deformation fields <= 1e-4
relative L2; applied SDFs <= 1e-3 cells max-abs.  Most kernels reproduce the reference's fp32
operation order and are in fact compared bit-exactly here; the CG differs only in the order of
its fp64 reductions."""
import numpy as np
import pytest

from conftest import rel_l2, sdf_pair
from oracle import port

pytestmark = pytest.mark.gpu

D = (14, 12, 13, 16)
SH = (D[3], D[2], D[1], D[0])
DEFO_TOL = 1e-4   # relative L2, deformation fields
SDF_TOL = 1e-3    # max-abs in cells, applied SDFs


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


@pytest.mark.parametrize("elem", [1, 4])
def test_interpolate_grid4d(gpu, elem):
    src = rnd(SH + ((4,) if elem == 4 else ()), 1)
    for tdims, off, sc in [((20, 9, 13, 11), 0., 1.), ((7, 6, 6, 8), (1.4, 1.2, 1.3, 3.0), 0.8),
                           ((28, 24, 26, 32), 0., (1., 1., 1., 0.9))]:
        eq(gpu.interpolate_grid4d(src, tdims, off, sc), port.interpolate_grid4d(src, tdims, off, sc))
    down = (D[0] // 2, D[1] // 2, D[2] // 2, D[3] // 2)
    a = gpu.interpol_grid_templ(src, down)
    eq(a, port.interpol_grid_templ(src, down))
    eq(gpu.interpol_grid_templ(a, D), port.interpol_grid_templ(a, D))


@pytest.mark.parametrize("elem", [1, 4])
def test_advect4d(gpu, elem):
    vel = rnd(SH + (4,), 3, 2.5)
    g = rnd(SH + ((4,) if elem == 4 else ()), 4)
    eq(gpu.advect4d(vel, g, 0.7), port.advect4d(vel, g, 0.7))
    eq(gpu.advect_cfl4d(1.5, vel, g, 1.0), port.advect_cfl4d(1.5, vel, g, 1.0))
    eq(gpu.advect_cfl4d(999., vel, g, 0.5), port.advect_cfl4d(999., vel, g, 0.5))


def test_bounds_minmax_ops(gpu):
    a = rnd(SH, 5)
    v = rnd(SH + (4,), 6)
    for w in (0, 1, 3):
        eq(gpu.set_bound4d(a, 0.1, w), port.set_bound4d(a, 0.1, w))
        eq(gpu.set_bound4d(v, (1, 2, 3, 4), w), port.set_bound4d(v, (1, 2, 3, 4), w))
        eq(gpu.set_bound_neumann4d(a, w), port.set_bound_neumann4d(a, w))
        eq(gpu.set_bound_neumann4d(v, w), port.set_bound_neumann4d(v, w))
    assert gpu.min_max4d(a) == port.min_max4d(a)
    assert gpu.min_max4d(v) == port.min_max4d(v)
    for elem in (1, 4):
        sh = SH + ((4,) if elem == 4 else ())
        x, y = rnd(sh, 7), rnd(sh, 8)
        f = (0.3, -1.5, 2.0, 0.25) if elem == 4 else 0.3
        for op in ("add", "sub", "mult"):
            eq(gpu.grid_op4d(op, x, y), port.grid_op4d(op, x, y))
        eq(gpu.grid_op4d("addScaled", x, y, f), port.grid_op4d("addScaled", x, y, f))
        eq(gpu.grid_op4d("multConst", x, None, f), port.grid_op4d("multConst", x, None, f))
        eq(gpu.grid_op4d("addConst", x, None, f), port.grid_op4d("addConst", x, None, f))
        eq(gpu.grid_op4d("clamp", x, None, (-0.5, 0.25)), port.grid_op4d("clamp", x, None, (-0.5, 0.25)))
    # ragged length (not a multiple of 4 floats) exercises the scalar tail
    r = rnd((5, 3, 3, 3), 9)
    eq(gpu.grid_op4d("multConst", r, None, 1.5), port.grid_op4d("multConst", r, None, 1.5))


@pytest.mark.parametrize("sigma", [0.5, 1.0, 1.125, 2.0, 3.0, 5.0])
def test_gaussian_blur_bitexact(gpu, sigma):
    """half-widths 1 and 2 run the register-tiled kernels, 3 / 4 the generic one, 5 (postVelBlur 10) its run-time-width form"""
    v = rnd(SH + (4,), 9)
    eq(gpu.gaussian_blur4d(v, sigma), port.gaussian_blur4d(v, sigma))
    if sigma >= 3.0:
        s1 = rnd(SH, 10)
        eq(gpu.gaussian_blur4d(s1, sigma), port.gaussian_blur4d(s1, sigma))


def test_separable_blur_option_is_close_but_not_the_default(gpu):
    """blur_mode 1 (opt-in): four 1D passes.  Same clipping / normalisation, fp32 rounding differs: ~1e-6 relative."""
    assert gpu.ctx.get_option("blur_mode") == 0
    v = rnd(SH + (4,), 9)
    s1 = rnd(SH, 10)
    try:
        gpu.ctx.set_option("blur_mode", 1)
        for sigma in (1.0, 2.0, 3.0):
            for x in (v, s1):
                got, want = gpu.gaussian_blur4d(x, sigma), port.gaussian_blur4d(x, sigma)
                assert rel_l2(got, want) < 2e-6
                assert not np.array_equal(got, want) or sigma == 1.0
                shell = np.ones(SH, bool)
                shell[1:-1, 1:-1, 1:-1, 1:-1] = False
                assert np.array_equal(got[shell], want[shell])      # the border shell is handled like the exact path
    finally:
        gpu.ctx.set_option("blur_mode", 0)


def test_optical_flow4d(gpu):
    i0, i1 = sdf_pair(D)
    v0 = np.zeros(SH + (4,), np.float32)
    a, rhs_a, it_a = gpu.optical_flow4d(v0, i0, i1, 1e-3, 1e-4, 0., 1e-2, -1., want_rhs=True, want_iters=True)
    b, rhs_b, it_b = port.optical_flow4d(v0, i0, i1, 1e-3, 1e-4, 0., 1e-2, -1., want_rhs=True, want_iters=True)
    eq(rhs_a, rhs_b)                      # assembly: bit-exact
    assert it_a == it_b                   # same stopping iteration as the reference
    assert rel_l2(a, b) <= DEFO_TOL
    # tight accuracy: many iterations, reduction-order effects stay at round-off level
    a, it_a = gpu.optical_flow4d(v0, i0, i1, 1e-3, 1e-4, 0., 1e-4, -1., want_iters=True)
    b, it_b = port.optical_flow4d(v0, i0, i1, 1e-3, 1e-4, 0., 1e-4, -1., want_iters=True)
    assert abs(it_a - it_b) <= 1 and rel_l2(a, b) <= DEFO_TOL
    # non-zero incoming velocity (smoothness / Tikhonov rhs terms), blur and border reset
    v1 = rnd(SH + (4,), 10, 0.3)
    a = gpu.optical_flow4d(v1, i0, i1, 1e-3, 1e-4, 4., 1e-2, 0.1)
    b = port.optical_flow4d(v1, i0, i1, 1e-3, 1e-4, 4., 1e-2, 0.1)
    assert rel_l2(a, b) <= DEFO_TOL
    # zero right-hand side: early out with zero iterations
    a, it = gpu.optical_flow4d(v0, i0, i0, 1e-3, 1e-4, 0., 1e-2, -1., want_iters=True)
    assert it == 0 and not a.any()


def test_project_and_expol_bitexact(gpu):
    i0, i1 = sdf_pair(D)
    vel = rnd(SH + (4,), 11, 0.7)
    da, ma = gpu.project_cells(vel, i0, i1, 4., 40)
    db, mb = port.project_cells(vel, i0, i1, 4., 40)
    eq(ma, mb)
    eq(da, db)
    eq(gpu.cv_expol_blur4d(da, ma, 5), port.cv_expol_blur4d(db, mb, 5))
    eq(gpu.cv_expol_blur4d(da, ma, 4), port.cv_expol_blur4d(db, mb, 4))
    z = np.zeros_like(vel)
    a = gpu.corr_vels_of4d(z, vel, i0, i1, 4., 4., 0.1, 40)
    b = port.corr_vels_of4d(z, vel, i0, i1, 4., 4., 0.1, 40)
    eq(a[0], b[0])
    eq(a[1], b[1])


@pytest.mark.parametrize("dims", [(16, 12, 10, 9), (8, 7, 6, 5), (40, 9, 5, 6), (12, 13, 7, 6)])
def test_expol_work_list_paths_bitexact(gpu, dims):
    """Random markers make ragged work lists (isolated cells, runs, row ends).  The default Vec4 work-list kernel is
    checked here; FLOF_EXPOL_MODE=0 in the environment runs the same cases through the component-plane kernel
    (4 x-cells per lane, edge values from the neighbour lanes), FLOF_EXPOL_MODE=2 through the dense kernel."""
    sh = (dims[3], dims[2], dims[1], dims[0])
    rng = np.random.default_rng(dims[0] * 131 + dims[1])
    for density, sweeps in ((0.25, 3), (0.7, 2), (0.02, 4), (1.0, 1)):
        a = rnd(sh + (4,), 21, 1.3)
        mark = (rng.random(sh) >= density).astype(np.float32)   # marker != 0 -> cell keeps its value
        want = port.cv_expol_blur4d(a, mark, sweeps)
        eq(gpu.cv_expol_blur4d(a, mark, sweeps), want)
        default_mode = gpu.ctx.get_option("expol_mode")
        try:
            for mode in (1, 3, 4, 5, 6):                              # 4y / 4y x 2z / 4y x 4z items (odd nz, clamped planes, ragged lists)
                gpu.ctx.set_option("expol_mode", mode)
                eq(gpu.cv_expol_blur4d(a, mark, sweeps), want)
        finally:
            gpu.ctx.set_option("expol_mode", default_mode)


def test_calc_ls_diff(gpu):
    i0, i1 = sdf_pair(D)
    for bnd in (0, 2):
        ra, oa = gpu.calc_ls_diff4d(i0, i1, 0.005, bnd, want_out=True)
        rb, ob = port.calc_ls_diff4d(i0, i1, 0.005, bnd, want_out=True)
        assert abs(ra - rb) <= 1e-6 * abs(rb)
        eq(oa, ob)


def test_calc_smoke_diff(gpu):
    """calcSmokeDiff4d (ref optflow4d.cpp:2140-2168): sum of float(|i0 - i1| * correction) over the cells inside bnd,
    accumulated in double, * 1e6 / cells.  The kernel sums in a different (deterministic) order: last bits of the fp64
    sum only."""
    i0, i1 = sdf_pair(D)
    for bnd, corr in ((0, 1.), (2, 200.)):
        sl = (slice(bnd, SH[0] - bnd), slice(bnd, SH[1] - bnd), slice(bnd, SH[2] - bnd), slice(bnd, SH[3] - bnd))
        d = (np.abs(i0[sl] - i1[sl]) * np.float32(corr)).astype(np.float32)
        want = d.astype(np.float64).sum() * 1e6 / d.size
        got = gpu.calc_smoke_diff4d(i0, i1, corr, bnd)
        assert abs(got - want) <= 2e-6 * abs(want), (got, want)


def test_host_buffer_entry_equals_resident_call(gpu):
    """flof_optical_flow_multiscale4d_host (the end-to-end plugin call: host buffers in and out) delivers the bits of the
    resident call."""
    from ofblend_b200 import capi
    d = (24, 24, 24, 20)
    i0, i1 = sdf_pair(d, seed=3)
    v0 = np.zeros((d[3], d[2], d[1], d[0], 4), np.float32)
    kw = dict(wSmooth=1e-3, wEnergy=1e-4, postVelBlur=4., cgAccuracy=1e-2, cfl=999., resetBndWidth=0.1,
              multiStep=3, minGridSize=20, doFinalProject=True)
    a, it_a, err_a = gpu.optical_flow_multiscale4d(v0, i0, i1, want_trace=True, **kw)
    vh = np.ascontiguousarray(v0.copy())
    err_b, tr = gpu.ctx.optical_flow_multiscale4d_host(vh, np.ascontiguousarray(i0), np.ascontiguousarray(i1), capi.make_params(**kw))
    assert list(tr.cg_iters[:tr.n_solves]) == it_a
    assert np.array_equal(vh, a)
    assert err_b == err_a[-1]


def test_extrapolation_marker_bitexact(gpu):
    i0, _ = sdf_pair(D)
    phi = port.set_bound4d(i0 / np.float32(-0.005), 0.1, 1)
    for inside in (False, True):
        pa, ma = gpu.extrap4d_ls_simple(phi, 6, inside, want_marker=True)
        pb, mb = port.extrap4d_ls_simple(phi, 6, inside, want_marker=True)
        assert ma.dtype == np.int32 and np.array_equal(ma, mb)   # Grid4d<int>: bit-exact
        eq(pa, pb)
        eq(gpu.extrap4d_ls_simple(phi, 40, inside), port.extrap4d_ls_simple(phi, 40, inside))
    vel = rnd(SH + (4,), 13)
    eq(gpu.extrapolate_vec4_simple(vel, phi, 5), port.extrapolate_vec4_simple(vel, phi, 5))
    a = rnd(SH, 14)
    eq(gpu.repeat_frame4d(a, 4.3, 3.0, 0), port.repeat_frame4d(a, 4.3, 3.0, 0))
    eq(gpu.repeat_frame4d(a, 5.0, 2.0, 1), port.repeat_frame4d(a, 5.0, 2.0, 1))


def test_3d_output_ops(gpu):
    a = rnd((13, 12, 14), 15)
    b = rnd((13, 12, 14), 16)
    eq(gpu.simple_blur_special(a, 1, 0., 2), port.simple_blur_special(a, 1, 0., 2))
    eq(gpu.simple_blur_special(a, 2, -999., 1), port.simple_blur_special(a, 2, -999., 1))
    eq(gpu.grid3_set_bound(a, 0.5, 2), port.grid3_set_bound(a, 0.5, 2))
    eq(gpu.levelset_join(a, b), port.levelset_join(a, b))


def test_mode3_loaders(gpu):
    nfiles, sx = 18, 16
    slices = rnd((nfiles, sx, sx, sx), 17)
    d = (20, 20, 20, 30)
    phi0 = np.zeros((d[3], d[2], d[1], d[0]), np.float32)
    off, sc = (2., 2., 2., 6.), (0.8, 0.8, 0.8, 0.8)
    kw = dict(fileIdxStart=0, fileIdxEnd=nfiles, spread=1., overrideSize=(20, 20, 20, 30), rescaleSdfValues=True,
              sdfIsoOff=-0.5, repeatStartFrame=0.1)
    a = gpu.load_place_grid4d(slices, phi0, off, sc, **kw)
    b = port.load_place_grid4d(slices, phi0, off, sc, **kw)
    eq(a, b)
    eq(gpu.load_place_grid4d(slices, a, off, sc, overrideGoodRegion=22, overrideTimeOff=-3., **kw),
       port.load_place_grid4d(slices, b, off, sc, overrideGoodRegion=22, overrideTimeOff=-3., **kw))
    eq(gpu.shift_forw_grid4d(a, 26), port.shift_forw_grid4d(b, 26))
    dd = (8, 8, 8, 12)
    defo = rnd((dd[3], dd[2], dd[1], dd[0], 4), 18, 0.4)
    big = (44, 44, 44, 30)
    phi = rnd((big[3], big[2], big[1], big[0]), 19)
    fac = tuple(big[i] / dd[i] for i in range(4))
    for tm in (6.5, 7.5, 12.5):
        eq(gpu.load_advect_time_slice(defo, big[:3], phi, tm, 0.5, 1., 0., 1., fac, big, 0., 4, 1.),
           port.load_advect_time_slice(defo, big[:3], phi, tm, 0.5, 1., 0., 1., fac, big, 0., 4, 1.))


@pytest.mark.parametrize("case", ["two", "two_aligned", "three", "two_eof"])
def test_defo_volumes_bitexact(gpu, case):
    """`thirdload`: two / three deformation volumes composed per output frame -- window refresh (all three
    branches of updateDefoVol incl. the re-used file handle), the compositions of ref optflow4d.cpp:2015-2089 and the slice
    look-up, through flof_defovol_window_update / flof_defovol_compose / flof_lookup_slice4d_with_vel; bit-exact against
    the oracle (which tests/test_oracle_vs_ref.py pins to the reference on the same cases)."""
    dd = (8, 7, 8, 20)
    vols = [rnd((dd[3], dd[2], dd[1], dd[0], 4), 30 + q, 0.5) for q in range(3 if case == "three" else 2)]
    big = (30, 28, 30, 30)
    phi = rnd((big[3], big[2], big[1], big[0]), 41)
    fac = tuple(big[i] / dd[i] for i in range(4))
    times = [0.9, 1.2, 2.4, 3.6, 3.7, 14.5, 16., 29.]
    if case == "two_eof":
        times = [24.55, 26.2, 27.85, 29.05]
    kw = dict(doAligned=(case == "two_aligned"), partialLoadFac=0.1, overrideSize=big, overrideTimeOff=0.5, bordSkip=3,
              defoAniFac=0.75)
    a = gpu.load_advect_defovols(vols, big[:3], phi, times, 0.6, 0.3, 0.2, 1., 0., 1., fac, **kw)
    b = port.load_advect_defovols(vols, big[:3], phi, times, 0.6, 0.3, 0.2, 1., 0., 1., fac, **kw)
    eq(a, b)
    assert np.abs(b).max() > 0


def test_unoptimised_load_advect_time_slice(gpu):
    """loadAdvectTimeSlice, the slower twin (ref optflow4d.cpp:1671-1760): dst over the interior cells (bnd 1) and the
    debugVel / debugVelT outputs, bit-exact against the oracle (pinned to the reference in tests/test_oracle_vs_ref.py)."""
    dd = (8, 9, 7, 12)
    defo = rnd((dd[3], dd[2], dd[1], dd[0], 4), 51, 0.6)
    big = (21, 19, 20, 30)
    phi = rnd((big[3], big[2], big[1], big[0]), 52)
    fac = tuple(big[i] / dd[i] for i in range(4))
    for tm, zero in ((6.5, False), (11.25, False), (29., False), (8., True)):
        a = gpu.load_advect_time_slice_unopt(defo, big[:3], phi, tm, 0.7, 1., 0., 1., fac, big, -0.5, 0.8, zero)
        b = port.load_advect_time_slice_unopt(defo, big[:3], phi, tm, 0.7, 1., 0., 1., fac, big, -0.5, 0.8, zero)
        for x, y in zip(a, b):
            eq(x, y)


def test_multiscale_small(gpu):
    """Full V-cycle (2 levels, 3 steps, final projection) against the oracle."""
    d = (24, 24, 24, 20)
    i0, i1 = sdf_pair(d, seed=3)
    v0 = np.zeros((d[3], d[2], d[1], d[0], 4), np.float32)
    kw = dict(wSmooth=1e-3, wEnergy=1e-4, postVelBlur=4., cgAccuracy=1e-2, cfl=999., resetBndWidth=0.1,
              multiStep=3, minGridSize=20, doFinalProject=True)
    a, it_a, err_a = gpu.optical_flow_multiscale4d(v0, i0, i1, want_trace=True, **kw)
    b, it_b, err_b = port.optical_flow_multiscale4d(v0, i0, i1, want_trace=True, **kw)
    assert it_a == it_b, (it_a, it_b)     # CG stops at the reference's iteration in every solve
    assert np.allclose(err_a, err_b, rtol=1e-4)
    assert rel_l2(a, b) <= DEFO_TOL
    # applied deformation: advected SDF within 1e-3 cells (SDF is scaled by 0.005 -> cells)
    adv_a = gpu.advect4d(a, i0)
    adv_b = port.advect4d(b, i0)
    assert np.abs(adv_a - adv_b).max() / 0.005 <= SDF_TOL


def test_tree_dot_mode_stays_inside_the_references_own_band(gpu):
    """dot_mode 0 (tree reductions, the faster non-default option): same CG stopping iterations, pre-projection field
    within the 1e-4 bar; the default (dot_mode 1, reference summation order) is held to identical bits in
    tests/test_gpu_seqsum.py."""
    from ofblend_b200 import synth
    dims = (32, 32, 32, 24)
    i0 = synth.post_process(synth.two_drop_phi(dims, 0), port)
    i1 = synth.post_process(synth.two_drop_phi(dims, 1), port)
    v0 = np.zeros(i0.shape + (4,), np.float32)
    kw = dict(synth.MODE1_PARAMS)
    kw["doFinalProject"] = False
    gpu.ctx.set_option("dot_mode", 0)
    try:
        a, it_a, err_a = gpu.optical_flow_multiscale4d(v0, i0, i1, want_trace=True, **kw)
    finally:
        gpu.ctx.set_option("dot_mode", 1)
    b, it_b, err_b = port.optical_flow_multiscale4d(v0, i0, i1, want_trace=True, **kw)
    assert it_a == it_b
    assert rel_l2(a, b) <= DEFO_TOL
    assert np.allclose(err_a, err_b, rtol=1e-5)
