"""ctypes access to oracle/_ref/libofref.so -- the UNMODIFIED reference (thunil/ofblend)
compiled by oracle/Makefile (`make ref`) with our C-ABI glue oracle/ref_shim.cpp.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's reference/
cpu_baseline legs may import this module.  It is the strongest parity checker we have:
it *is* the reference's CPU implementation.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libofref.so")

_lib = None


def available():
    return os.path.isfile(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libofref.so missing: run `make -C oracle ref` "
                               "(needs /root/reference)")
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_last_error.restype = C.c_char_p
        _lib.ref_set_debug_level(0)
    return _lib


def _chk(rc):
    if rc != 0:
        raise RuntimeError("reference raised: " + lib().ref_last_error().decode())


def _i4(d):
    return (C.c_int * len(d))(*[int(x) for x in d])


def _f4(v):
    v = np.broadcast_to(np.asarray(v, dtype=np.float32), (4,))
    return (C.c_float * 4)(*[float(x) for x in v])


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def dims_of(a, elem=1):
    """numpy layout is [t, z, y, x(, 4)]; returns (nx, ny, nz, nt)."""
    s = a.shape[:4]
    return (s[3], s[2], s[1], s[0])


def set_threads(n):
    return lib().ref_set_threads(int(n))


def set_debug_level(l):
    lib().ref_set_debug_level(int(l))


def interpolate_grid4d(src, tdims, offset=0., scale=1., size=-1.):
    src = _f32(src)
    elem = 4 if src.ndim == 5 else 1
    sd = dims_of(src)
    shape = (tdims[3], tdims[2], tdims[1], tdims[0]) + ((4,) if elem == 4 else ())
    dst = np.zeros(shape, np.float32)
    _chk(lib().ref_interpolate_grid4d(_i4(sd), _p(src), _i4(tdims), _p(dst), elem,
                                      _f4(offset), _f4(scale), _f4(size)))
    return dst


def interpol_grid_templ(src, tdims):
    src = _f32(src)
    elem = 4 if src.ndim == 5 else 1
    shape = (tdims[3], tdims[2], tdims[1], tdims[0]) + ((4,) if elem == 4 else ())
    dst = np.zeros(shape, np.float32)
    _chk(lib().ref_interpol_grid_templ(_i4(dims_of(src)), _p(src), _i4(tdims), _p(dst), elem))
    return dst


def advect4d(vel, grid, dtFac=1., dt=1.):
    vel = _f32(vel)
    g = _f32(grid).copy()
    elem = 4 if g.ndim == 5 else 1
    _chk(lib().ref_advect4d(_i4(dims_of(vel)), _p(vel), _p(g), elem, C.c_float(dtFac),
                            C.c_float(dt)))
    return g


def advect_cfl4d(cfl, vel, grid, velFactor=1.):
    vel = _f32(vel)
    g = _f32(grid).copy()
    elem = 4 if g.ndim == 5 else 1
    _chk(lib().ref_advect_cfl4d(_i4(dims_of(vel)), C.c_float(cfl), _p(vel), _p(g), elem,
                                C.c_float(velFactor)))
    return g


def optical_flow4d(vel, i0, i1, wSmooth=0., wEnergy=0., postVelBlur=0., cgAccuracy=1e-4,
                   resetBndWidth=-1., want_rhs=False):
    v = _f32(vel).copy()
    i0 = _f32(i0)
    i1 = _f32(i1)
    rhs = np.zeros_like(i0) if want_rhs else None
    _chk(lib().ref_optical_flow4d(_i4(dims_of(i0)), _p(v), _p(i0), _p(i1), _p(rhs),
                                  C.c_float(wSmooth), C.c_float(wEnergy),
                                  C.c_float(postVelBlur), C.c_float(cgAccuracy), 1,
                                  C.c_float(resetBndWidth)))
    return (v, rhs) if want_rhs else v


def gaussian_blur4d(a, sigma, iters=1):
    g = _f32(a).copy()
    elem = 4 if g.ndim == 5 else 1
    _chk(lib().ref_gaussian_blur4d(_i4(dims_of(g)), _p(g), elem, C.c_float(sigma), int(iters)))
    return g


def project_cells(vel, phiOrg, phiTarget, threshPhi, maxIter):
    vel = _f32(vel)
    po = _f32(phiOrg)
    pt = _f32(phiTarget)
    dst = np.zeros_like(vel)
    marker = np.zeros_like(po)
    _chk(lib().ref_project_cells(_i4(dims_of(po)), _p(dst), _p(vel), _p(po), _p(pt),
                                 _p(marker), C.c_float(threshPhi), int(maxIter)))
    return dst, marker


def cv_expol_blur4d(a, marker, sweeps):
    g = _f32(a).copy()
    mk = _f32(marker)
    _chk(lib().ref_cv_expol_blur4d(_i4(dims_of(mk)), _p(g), _p(mk), int(sweeps)))
    return g


def corr_vels_of4d(dst, vel, phiOrg, phiTarget, threshPhi=1e10, postVelBlur=0.,
                   resetBndWidth=-1., maxIter=100):
    d = _f32(dst).copy()
    v = _f32(vel).copy()
    po = _f32(phiOrg)
    pt = _f32(phiTarget)
    _chk(lib().ref_corr_vels_of4d(_i4(dims_of(po)), _p(d), _p(v), _p(po), _p(pt),
                                  C.c_float(threshPhi), C.c_float(postVelBlur),
                                  C.c_float(resetBndWidth), int(maxIter)))
    return d, v


def calc_ls_diff4d(i0, i1, correction=1., bnd=0, want_out=False):
    i0 = _f32(i0)
    i1 = _f32(i1)
    out = np.zeros_like(i0) if want_out else None
    res = C.c_float(0)
    _chk(lib().ref_calc_ls_diff4d(_i4(dims_of(i0)), _p(i0), _p(i1), _p(out),
                                  C.c_float(correction), int(bnd), C.byref(res)))
    return (res.value, out) if want_out else res.value


def optical_flow_multiscale4d(vel, i0, i1, wSmooth=0., wEnergy=0., postVelBlur=0.,
                              cgAccuracy=1e-4, cfl=999., resetBndWidth=-1., multiStep=1,
                              projSizeThresh=9999, minGridSize=10, doFinalProject=False):
    v = _f32(vel).copy()
    i0 = _f32(i0)
    i1 = _f32(i1)
    _chk(lib().ref_optical_flow_multiscale4d(
        _i4(dims_of(i0)), _p(v), _p(i0), _p(i1), C.c_float(wSmooth), C.c_float(wEnergy),
        C.c_float(postVelBlur), C.c_float(cgAccuracy), C.c_float(cfl),
        C.c_float(resetBndWidth), int(multiStep), int(projSizeThresh), int(minGridSize),
        int(bool(doFinalProject))))
    return v


# ---- 3D instantiations: numpy layout [z, y, x(, 3)]
def _d3v(a):
    return (a.shape[2], a.shape[1], a.shape[0])


def optical_flow_multiscale3d(vel, i0, i1, wSmooth=0., wEnergy=0., postVelBlur=0., cgAccuracy=1e-4, cfl=999.,
                              resetBndWidth=-1., multiStep=1, projSizeThresh=9999, minGridSize=10, doFinalProject=False):
    v = _f32(vel).copy()
    i0 = _f32(i0)
    i1 = _f32(i1)
    _chk(lib().ref_optical_flow_multiscale3d(
        _i4(_d3v(i0)), _p(v), _p(i0), _p(i1), C.c_float(wSmooth), C.c_float(wEnergy), C.c_float(postVelBlur),
        C.c_float(cgAccuracy), C.c_float(cfl), C.c_float(resetBndWidth), int(multiStep), int(projSizeThresh),
        int(minGridSize), int(bool(doFinalProject))))
    return v


def corr_vels_of3d(dst, vel, phiOrg, phiTarget, threshPhi=1e10, postVelBlur=0., resetBndWidth=-1., maxIter=100):
    d = _f32(dst).copy()
    v = _f32(vel).copy()
    po = _f32(phiOrg)
    pt = _f32(phiTarget)
    _chk(lib().ref_corr_vels_of3d(_i4(_d3v(po)), _p(d), _p(v), _p(po), _p(pt), C.c_float(threshPhi),
                                  C.c_float(postVelBlur), C.c_float(resetBndWidth), int(maxIter)))
    return d, v


def advect_semi_lagrange_cfl3d(cfl, vel, grid, velFactor=1.):
    v = _f32(vel)
    g = _f32(grid).copy()
    _chk(lib().ref_advect_semi_lagrange_cfl3d(_i4(_d3v(v)), C.c_float(cfl), _p(v), _p(g), 3 if g.ndim == 4 else 1,
                                              C.c_float(velFactor)))
    return g


def calc_ls_diff3d(i0, i1, correction=1., bnd=0, want_out=False):
    a = _f32(i0)
    b = _f32(i1)
    out = np.zeros(a.shape, np.float32) if want_out else None
    r = C.c_float(0)
    _chk(lib().ref_calc_ls_diff3d(_i4(_d3v(a)), _p(a), _p(b), _p(out), C.c_float(correction), int(bnd), C.byref(r)))
    return (r.value, out) if want_out else r.value


def extrap4d_ls_simple(phi, distance=4, inside=False, want_marker=False):
    p = _f32(phi).copy()
    if want_marker:
        mk = np.zeros(p.shape, np.int32)
        _chk(lib().ref_extrap4d_ls_simple_marker(_i4(dims_of(p)), _p(p), int(distance),
                                                 int(bool(inside)), _p(mk)))
        return p, mk
    _chk(lib().ref_extrap4d_ls_simple(_i4(dims_of(p)), _p(p), int(distance), int(bool(inside))))
    return p


def extrapolate_vec4_simple(vel, phi, distance):
    v = _f32(vel).copy()
    p = _f32(phi)
    _chk(lib().ref_extrapolate_vec4_simple(_i4(dims_of(p)), _p(v), _p(p), int(distance)))
    return v


def repeat_frame4d(phi, srct, rng=0., bnd=0):
    p = _f32(phi).copy()
    _chk(lib().ref_repeat_frame4d(_i4(dims_of(p)), _p(p), C.c_float(srct), C.c_float(rng),
                                  int(bnd)))
    return p


def set_bound4d(a, value, w=1):
    if a.dtype == np.int32:
        g = np.ascontiguousarray(a).copy()
        elem = -1
    else:
        g = _f32(a).copy()
        elem = 4 if g.ndim == 5 else 1
    _chk(lib().ref_set_bound4d(_i4(dims_of(g)), _p(g), elem, _f4(value), int(w)))
    return g


def set_bound_neumann4d(a, w=1):
    g = _f32(a).copy()
    elem = 4 if g.ndim == 5 else 1
    _chk(lib().ref_set_bound_neumann4d(_i4(dims_of(g)), _p(g), elem, int(w)))
    return g


def min_max4d(a):
    g = _f32(a)
    elem = 4 if g.ndim == 5 else 1
    out = np.zeros(3, np.float32)
    _chk(lib().ref_min_max4d(_i4(dims_of(g)), _p(g), elem, _p(out)))
    return tuple(float(x) for x in out)


GRID_OPS = {"add": 0, "sub": 1, "mult": 2, "addScaled": 3, "multConst": 4, "addConst": 5,
            "clamp": 6}


def grid_op4d(op, a, b=None, factor=0.):
    g = _f32(a).copy()
    elem = 4 if g.ndim == 5 else 1
    bb = _f32(b) if b is not None else None
    fac = np.zeros(4, np.float32)
    f = np.atleast_1d(np.asarray(factor, np.float32))
    if op == "clamp":
        fac[:2] = f[:2]
    else:
        fac[:] = np.broadcast_to(f, (4,)) if f.size in (1, 4) else 0
    _chk(lib().ref_grid_op4d(_i4(dims_of(g)), _p(g), _p(bb), elem, GRID_OPS[op],
                             (C.c_float * 4)(*[float(x) for x in fac])))
    return g


def mult_const(a, s):
    return grid_op4d("multConst", a, None, s)


def _d3(a):
    return (a.shape[2], a.shape[1], a.shape[0])


def simple_blur_special(a, iters=1, thresh=0., bord=0):
    g = _f32(a).copy()
    _chk(lib().ref_simple_blur_special(_i4(_d3(g)), _p(g), int(iters), C.c_float(thresh),
                                       int(bord)))
    return g


def grid3_set_bound(a, value, w=1):
    g = _f32(a).copy()
    _chk(lib().ref_grid3_set_bound(_i4(_d3(g)), _p(g), C.c_float(value), int(w)))
    return g


def levelset_join(a, b):
    g = _f32(a).copy()
    _chk(lib().ref_levelset_join(_i4(_d3(g)), _p(g), _p(_f32(b))))
    return g


def init_test_checkerboard(dims, brd=0, want_vec=False):
    val = np.zeros((dims[3], dims[2], dims[1], dims[0]), np.float32)
    vec = np.zeros(val.shape + (4,), np.float32) if want_vec else None
    _chk(lib().ref_init_test_checkerboard(_i4(dims), _p(val), _p(vec), int(brd)))
    return (val, vec) if want_vec else val


def grid4d_save(a, name):
    g = _f32(a)
    elem = 4 if g.ndim == 5 else 1
    _chk(lib().ref_grid4d_save(_i4(dims_of(g)), _p(g), elem, name.encode()))


def grid4d_load(dims, elem, name):
    shape = (dims[3], dims[2], dims[1], dims[0]) + ((4,) if elem == 4 else ())
    g = np.zeros(shape, np.float32)
    _chk(lib().ref_grid4d_load(_i4(dims), _p(g), elem, name.encode()))
    return g


def grid3_save(a, name):
    g = _f32(a)
    _chk(lib().ref_grid3_save(_i4(_d3(g)), _p(g), name.encode()))


def grid3_load(d3, name):
    g = np.zeros((d3[2], d3[1], d3[0]), np.float32)
    _chk(lib().ref_grid3_load(_i4(d3), _p(g), name.encode()))
    return g


def load_place_grid4d(fname, phi, offset, scale, fileIdxStart=-1, fileIdxEnd=-1,
                      debugSkipLoad=999999, spread=1., overrideSize=-1., overrideTimeOff=0.,
                      overrideGoodRegion=0, loadTimeScale=1., rescaleSdfValues=False,
                      sdfIsoOff=0., repeatStartFrame=0.):
    p = _f32(phi).copy()
    _chk(lib().ref_load_place_grid4d(
        fname.encode(), _i4(dims_of(p)), _p(p), _f4(offset), _f4(scale), int(fileIdxStart),
        int(fileIdxEnd), int(debugSkipLoad), C.c_float(spread), _f4(overrideSize),
        C.c_float(overrideTimeOff), int(overrideGoodRegion), C.c_float(loadTimeScale),
        int(bool(rescaleSdfValues)), C.c_float(sdfIsoOff), C.c_float(repeatStartFrame)))
    return p


def shift_forw_grid4d(phi, overrideGoodRegion):
    p = _f32(phi).copy()
    _chk(lib().ref_shift_forw_grid4d(_i4(dims_of(p)), _p(p), int(overrideGoodRegion)))
    return p


def load_advect_time_slices_opt(fname, d3, phi, times, blendAlpha, loadTimeScale, defoOffset,
                                defoScale, defoFactor, overrideSize=-1., overrideTimeOff=0.,
                                bordSkip=1, defoAniFac=1.):
    p = _f32(phi)
    times = np.ascontiguousarray(times, np.float32)
    out = np.zeros((len(times), d3[2], d3[1], d3[0]), np.float32)
    _chk(lib().ref_load_advect_time_slices_opt(
        fname.encode(), _i4(d3), _p(out), _i4(dims_of(p)), _p(p), len(times), _p(times),
        C.c_float(blendAlpha), C.c_float(loadTimeScale), _f4(defoOffset), _f4(defoScale),
        _f4(defoFactor), _f4(overrideSize), C.c_float(overrideTimeOff), int(bordSkip),
        C.c_float(defoAniFac)))
    return out


def load_advect_defovols(fnames, d3, phi, times, blendAlpha, thirdAlpha, fourthAlpha, loadTimeScale, defoOffset, defoScale,
                         defoFactor, doAligned=False, partialLoadFac=0.2, overrideSize=-1., overrideTimeOff=0.,
                         bordSkip=1, defoAniFac=1.):
    """The reference's _OptInit(useDefoVols=True) + _OptAdd per further file + n x _OptRun + _Finish; fnames: .uni files."""
    p = _f32(phi)
    times = np.ascontiguousarray(times, np.float32)
    out = np.zeros((len(times), d3[2], d3[1], d3[0]), np.float32)
    names = (C.c_char_p * 3)(*([f.encode() for f in fnames] + [None] * (3 - len(fnames))))
    _chk(lib().ref_load_advect_defovols(
        names, len(fnames), int(bool(doAligned)), C.c_float(partialLoadFac), _i4(d3), _p(out), _i4(dims_of(p)), _p(p),
        len(times), _p(times), C.c_float(blendAlpha), C.c_float(thirdAlpha), C.c_float(fourthAlpha),
        C.c_float(loadTimeScale), _f4(defoOffset), _f4(defoScale), _f4(defoFactor), _f4(overrideSize),
        C.c_float(overrideTimeOff), int(bordSkip), C.c_float(defoAniFac)))
    return out


def load_advect_time_slice_unopt(fname, d3, phi, time, blendAlpha, loadTimeScale, defoOffset, defoScale, defoFactor,
                                 overrideSize=-1., overrideTimeOff=0., defoAniFac=1., zeroVel=False):
    p = _f32(phi)
    out = np.zeros((d3[2], d3[1], d3[0]), np.float32)
    dv = np.zeros((d3[2], d3[1], d3[0], 3), np.float32)
    dt = np.zeros((d3[2], d3[1], d3[0]), np.float32)
    _chk(lib().ref_load_advect_time_slice_debug(
        fname.encode(), _i4(d3), _p(out), _p(dv), _p(dt), _i4(dims_of(p)), _p(p), C.c_float(time), C.c_float(blendAlpha),
        C.c_float(loadTimeScale), _f4(defoOffset), _f4(defoScale), _f4(defoFactor), _f4(overrideSize),
        C.c_float(overrideTimeOff), C.c_float(defoAniFac), int(bool(zeroVel))))
    return out, dv, dt


def load_advect_time_slice(fname, d3, phi, time, blendAlpha, loadTimeScale, defoOffset,
                           defoScale, defoFactor, overrideSize=-1., overrideTimeOff=0.):
    p = _f32(phi)
    out = np.zeros((d3[2], d3[1], d3[0]), np.float32)
    _chk(lib().ref_load_advect_time_slice(
        fname.encode(), _i4(d3), _p(out), _i4(dims_of(p)), _p(p), C.c_float(time),
        C.c_float(blendAlpha), C.c_float(loadTimeScale), _f4(defoOffset), _f4(defoScale),
        _f4(defoFactor), _f4(overrideSize), C.c_float(overrideTimeOff)))
    return out


def get_slice_from4d(src, srct):
    s = _f32(src)
    out = np.zeros(s.shape[1:4], np.float32)
    _chk(lib().ref_get_slice_from4d(_i4(dims_of(s)), _p(s), int(srct), _p(out)))
    return out
