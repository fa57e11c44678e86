"""ctypes access to oracle/libflof_oracle.so -- our plain-C restatement of the reference
(oracle/flof_oracle.c).  Same Python surface as oracle/ref.py so tests can run either.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's reference/
cpu_baseline legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libflof_oracle.so")
_lib = None


class Dim4(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("nt", C.c_int)]


def build():
    """Compile the C restatement (gcc only; works on any box)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH) or (
                os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "flof_oracle.c"))):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_calc_ls_diff4d.restype = C.c_float
        _lib.orc_optical_flow_multiscale4d.restype = C.c_float
        _lib.orc_dot_seq.restype = C.c_double
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f4(v):
    v = np.broadcast_to(np.asarray(v, dtype=np.float32), (4,))
    return (C.c_float * 4)(*[float(x) for x in v])


def dims_of(a):
    s = a.shape[:4]
    return (s[3], s[2], s[1], s[0])


def _d(dims):
    return Dim4(*[int(x) for x in dims])


def _shape(dims, elem):
    return (dims[3], dims[2], dims[1], dims[0]) + ((4,) if elem == 4 else ())


def _elem(a):
    return 4 if a.ndim == 5 else 1


def set_threads(n):
    return lib().orc_set_threads(int(n))


def interpolate_grid4d(src, tdims, offset=0., scale=1., size=-1.):
    src = _f32(src)
    e = _elem(src)
    dst = np.zeros(_shape(tdims, e), np.float32)
    lib().orc_interpolate_grid4d(_p(dst), _d(tdims), _p(src), _d(dims_of(src)), e, _f4(offset),
                                 _f4(scale), _f4(size))
    return dst


def interpol_grid_templ(src, tdims):
    src = _f32(src)
    e = _elem(src)
    dst = np.zeros(_shape(tdims, e), np.float32)
    lib().orc_interpol_grid_templ(_p(dst), _d(tdims), _p(src), _d(dims_of(src)), e)
    return dst


def advect4d(vel, grid, dtFac=1., dt=1.):
    vel = _f32(vel)
    g = _f32(grid).copy()
    lib().orc_advect4d(_p(vel), _p(g), _d(dims_of(vel)), _elem(g),
                       C.c_float(np.float32(dt) * np.float32(dtFac)))
    return g


def advect_cfl4d(cfl, vel, grid, velFactor=1.):
    vel = _f32(vel)
    g = _f32(grid).copy()
    lib().orc_advect_cfl4d(C.c_float(cfl), _p(vel), _p(g), _d(dims_of(vel)), _elem(g),
                           C.c_float(velFactor))
    return g


def optical_flow4d(vel, i0, i1, wSmooth=0., wEnergy=0., postVelBlur=0., cgAccuracy=1e-4,
                   resetBndWidth=-1., want_rhs=False, want_iters=False):
    v = _f32(vel).copy()
    i0 = _f32(i0)
    i1 = _f32(i1)
    rhs = np.zeros_like(i0) if want_rhs else None
    res = C.c_float(0)
    it = lib().orc_optical_flow4d(_p(v), _p(i0), _p(i1), _p(rhs), _d(dims_of(i0)),
                                  C.c_float(wSmooth), C.c_float(wEnergy), C.c_float(postVelBlur),
                                  C.c_float(cgAccuracy), C.c_float(resetBndWidth), C.byref(res))
    out = (v,)
    if want_rhs:
        out += (rhs,)
    if want_iters:
        out += (it,)
    return out if len(out) > 1 else v


def gaussian_blur4d(a, sigma, iters=1):
    g = _f32(a).copy()
    lib().orc_gaussian_blur4d(_p(g), _d(dims_of(g)), _elem(g), C.c_float(sigma), int(iters))
    return g


def project_cells(vel, phiOrg, phiTarget, threshPhi, maxIter):
    vel = _f32(vel)
    po = _f32(phiOrg)
    pt = _f32(phiTarget)
    dst = np.zeros_like(vel)
    marker = np.zeros_like(po)
    lib().orc_project_cells(_p(dst), _p(vel), _p(po), _p(pt), _p(marker), _d(dims_of(po)),
                            C.c_float(threshPhi), int(maxIter))
    return dst, marker


def cv_expol_blur4d(a, marker, sweeps):
    g = _f32(a).copy()
    mk = _f32(marker)
    lib().orc_cv_expol_blur4d(_p(g), _p(mk), _d(dims_of(mk)), int(sweeps))
    return g


def corr_vels_of4d(dst, vel, phiOrg, phiTarget, threshPhi=1e10, postVelBlur=0.,
                   resetBndWidth=-1., maxIter=100):
    d = _f32(dst).copy()
    v = _f32(vel).copy()
    po = _f32(phiOrg)
    pt = _f32(phiTarget)
    lib().orc_corr_vels_of4d(_p(d), _p(v), _p(po), _p(pt), _d(dims_of(po)), C.c_float(threshPhi),
                             C.c_float(postVelBlur), C.c_float(resetBndWidth), int(maxIter))
    return d, v


def calc_ls_diff4d(i0, i1, correction=1., bnd=0, want_out=False):
    i0 = _f32(i0)
    i1 = _f32(i1)
    out = np.zeros_like(i0) if want_out else None
    r = lib().orc_calc_ls_diff4d(_p(i0), _p(i1), _p(out), _d(dims_of(i0)), C.c_float(correction),
                                 int(bnd))
    return (r, out) if want_out else r


def optical_flow_multiscale4d(vel, i0, i1, wSmooth=0., wEnergy=0., postVelBlur=0.,
                              cgAccuracy=1e-4, cfl=999., resetBndWidth=-1., multiStep=1,
                              projSizeThresh=9999, minGridSize=10, doFinalProject=False,
                              want_trace=False):
    v = _f32(vel).copy()
    i0 = _f32(i0)
    i1 = _f32(i1)
    iters = (C.c_int * 64)()
    errs = (C.c_float * 64)()
    ni = C.c_int(0)
    ne = C.c_int(0)
    lib().orc_optical_flow_multiscale4d(
        _p(v), _p(i0), _p(i1), _d(dims_of(i0)), C.c_float(wSmooth), C.c_float(wEnergy),
        C.c_float(postVelBlur), C.c_float(cgAccuracy), C.c_float(cfl), C.c_float(resetBndWidth),
        int(multiStep), int(projSizeThresh), int(minGridSize), int(bool(doFinalProject)), iters,
        C.byref(ni), errs, C.byref(ne))
    if want_trace:
        return v, list(iters[:ni.value]), [float(x) for x in errs[:ne.value]]
    return v


# ---- 3D / 2D instantiations (oracle/flof_oracle3.c): numpy layout [z, y, x(, 3)]
def _n3(a):
    return int(a.shape[2]), int(a.shape[1]), int(a.shape[0])


def advect_semi_lagrange_cfl3d(cfl, vel, grid, velFactor=1.):
    v = _f32(vel)
    g = _f32(grid).copy()
    lib().orc3_advect_cfl(C.c_float(cfl), _p(v), _p(g), 3 if g.ndim == 4 else 1, *_n3(v), C.c_float(velFactor))
    return g


def calc_ls_diff3d(i0, i1, correction=1., bnd=0, want_out=False):
    a = _f32(i0)
    b = _f32(i1)
    out = np.zeros(a.shape, np.float32) if want_out else None
    fn = lib().orc3_calc_ls_diff
    fn.restype = C.c_float
    r = fn(_p(a), _p(b), _p(out), *_n3(a), C.c_float(correction), int(bnd))
    return (r, out) if want_out else r


def corr_vels_of3d(dst, vel, phiOrg, phiTarget, threshPhi=1e10, postVelBlur=0., resetBndWidth=-1., maxIter=100):
    d = _f32(dst).copy()
    v = _f32(vel).copy()
    po = _f32(phiOrg)
    pt = _f32(phiTarget)
    lib().orc3_corr_vels(_p(d), _p(v), _p(po), _p(pt), *_n3(po), C.c_float(threshPhi), C.c_float(postVelBlur),
                         C.c_float(resetBndWidth), int(maxIter))
    return d, v


def optical_flow_multiscale3d(vel, i0, i1, wSmooth=0., wEnergy=0., postVelBlur=0., cgAccuracy=1e-4, cfl=999.,
                              resetBndWidth=-1., multiStep=1, projSizeThresh=9999, minGridSize=10, doFinalProject=False,
                              want_trace=False):
    v = _f32(vel).copy()
    i0 = _f32(i0)
    i1 = _f32(i1)
    iters = (C.c_int * 64)()
    errs = (C.c_float * 64)()
    ni = C.c_int(0)
    ne = C.c_int(0)
    fn = lib().orc3_optical_flow_multiscale
    fn.restype = C.c_float
    fn(_p(v), _p(i0), _p(i1), *_n3(i0), C.c_float(wSmooth), C.c_float(wEnergy), C.c_float(postVelBlur),
       C.c_float(cgAccuracy), C.c_float(cfl), C.c_float(resetBndWidth), int(multiStep), int(projSizeThresh),
       int(minGridSize), int(bool(doFinalProject)), iters, C.byref(ni), errs, C.byref(ne))
    if want_trace:
        return v, list(iters[:ni.value]), [float(x) for x in errs[:ne.value]]
    return v


def extrap4d_ls_simple(phi, distance=4, inside=False, want_marker=False):
    p = _f32(phi).copy()
    mk = np.zeros(p.shape, np.int32) if want_marker else None
    lib().orc_extrap4d_ls_simple(_p(p), _d(dims_of(p)), int(distance), int(bool(inside)), _p(mk))
    return (p, mk) if want_marker else p


def extrapolate_vec4_simple(vel, phi, distance):
    v = _f32(vel).copy()
    p = _f32(phi)
    lib().orc_extrapolate_vec4_simple(_p(v), _p(p), _d(dims_of(p)), int(distance))
    return v


def repeat_frame4d(phi, srct, rng=0., bnd=0):
    p = _f32(phi).copy()
    lib().orc_repeat_frame4d(_p(p), _d(dims_of(p)), C.c_float(srct), C.c_float(rng), int(bnd))
    return p


def set_bound4d(a, value, w=1):
    if a.dtype == np.int32:
        g = np.ascontiguousarray(a).copy()
        lib().orc_set_bound4d_int(_p(g), _d(dims_of(g)), int(np.atleast_1d(value)[0]), int(w))
        return g
    g = _f32(a).copy()
    lib().orc_set_bound4d(_p(g), _d(dims_of(g)), _elem(g), _f4(value), int(w))
    return g


def set_bound_neumann4d(a, w=1):
    g = _f32(a).copy()
    lib().orc_set_bound_neumann4d(_p(g), _d(dims_of(g)), _elem(g), int(w))
    return g


def min_max4d(a):
    g = _f32(a)
    out = np.zeros(3, np.float32)
    lib().orc_min_max4d(_p(g), _d(dims_of(g)), _elem(g), _p(out))
    return tuple(float(x) for x in out)


GRID_OPS = {"add": 0, "sub": 1, "mult": 2, "addScaled": 3, "multConst": 4, "addConst": 5,
            "clamp": 6}


def grid_op4d(op, a, b=None, factor=0.):
    g = _f32(a).copy()
    bb = _f32(b) if b is not None else None
    fac = np.zeros(4, np.float32)
    f = np.atleast_1d(np.asarray(factor, np.float32))
    if op == "clamp":
        fac[:2] = f[:2]
    else:
        fac[:] = np.broadcast_to(f, (4,))
    lib().orc_grid_op4d(_p(g), _p(bb), _d(dims_of(g)), _elem(g), GRID_OPS[op],
                        (C.c_float * 4)(*[float(x) for x in fac]))
    return g


def mult_const(a, s):
    return grid_op4d("multConst", a, None, s)


def simple_blur_special(a, iters=1, thresh=0., bord=0):
    g = _f32(a).copy()
    lib().orc_simple_blur_special(_p(g), g.shape[2], g.shape[1], g.shape[0], int(iters),
                                  C.c_float(thresh), int(bord))
    return g


def grid3_set_bound(a, value, w=1):
    g = _f32(a).copy()
    lib().orc_grid3_set_bound(_p(g), g.shape[2], g.shape[1], g.shape[0], C.c_float(value), int(w))
    return g


def levelset_join(a, b):
    g = _f32(a).copy()
    lib().orc_levelset_join(_p(g), _p(_f32(b)), C.c_long(g.size))
    return g


def load_place_grid4d(slices, phi, offset, scale, fileIdxStart=-1, fileIdxEnd=-1,
                      debugSkipLoad=999999, spread=1., overrideSize=-1., overrideTimeOff=0.,
                      overrideGoodRegion=0, loadTimeScale=1., rescaleSdfValues=False,
                      sdfIsoOff=0., repeatStartFrame=0.):
    """slices: float32 [nfiles, sz, sy, sx], slices[i] = file (fileIdxStart + i)."""
    s = _f32(slices)
    p = _f32(phi).copy()
    lib().orc_load_place_grid4d(
        _p(s), s.shape[3], s.shape[2], s.shape[1], _p(p), _d(dims_of(p)), _f4(offset), _f4(scale),
        int(fileIdxStart), int(fileIdxEnd), int(debugSkipLoad), C.c_float(spread),
        _f4(overrideSize), C.c_float(overrideTimeOff), int(overrideGoodRegion),
        C.c_float(loadTimeScale), int(bool(rescaleSdfValues)), C.c_float(sdfIsoOff),
        C.c_float(repeatStartFrame))
    return p


def shift_forw_grid4d(phi, overrideGoodRegion):
    p = _f32(phi).copy()
    lib().orc_shift_forw_grid4d(_p(p), _d(dims_of(p)), int(overrideGoodRegion))
    return p


def load_advect_time_slice(defo, d3, phi, time, blendAlpha, loadTimeScale, defoOffset, defoScale,
                           defoFactor, overrideSize=-1., overrideTimeOff=0., bordSkip=1,
                           defoAniFac=1., dst=None):
    defo = _f32(defo)
    p = _f32(phi)
    out = np.zeros((d3[2], d3[1], d3[0]), np.float32) if dst is None else _f32(dst).copy()
    lib().orc_load_advect_time_slice(
        _p(defo), _d(dims_of(defo)), _p(out), int(d3[0]), int(d3[1]), int(d3[2]), _p(p),
        _d(dims_of(p)), C.c_float(time), C.c_float(blendAlpha), C.c_float(loadTimeScale),
        _f4(defoOffset), _f4(defoScale), _f4(defoFactor), _f4(overrideSize),
        C.c_float(overrideTimeOff), int(bordSkip), C.c_float(defoAniFac))
    return out


def load_advect_defovols(vols, d3, phi, times, blendAlpha, thirdAlpha, fourthAlpha, loadTimeScale, defoOffset, defoScale,
                         defoFactor, doAligned=False, partialLoadFac=0.2, overrideSize=-1., overrideTimeOff=0.,
                         bordSkip=1, defoAniFac=1.):
    """loadAdvectTimeSlice_OptInit(useDefoVols=True) + _OptAdd + n x _OptRun (ref optflow4d.cpp:1822-1863, 1951-2105);
    vols: 2 or 3 complete deformation volumes (Vec4)."""
    vols = [_f32(v) for v in vols]
    p = _f32(phi)
    times = np.ascontiguousarray(times, np.float32)
    out = np.zeros((len(times), d3[2], d3[1], d3[0]), np.float32)
    arr = (C.c_void_p * 3)(*([v.ctypes.data for v in vols] + [None] * (3 - len(vols))))
    lib().orc_load_advect_defovols(
        arr, len(vols), _d(dims_of(vols[0])), int(bool(doAligned)), C.c_float(partialLoadFac), _p(out), int(d3[0]), int(d3[1]),
        int(d3[2]), _p(p), _d(dims_of(p)), len(times), _p(times), C.c_float(blendAlpha), C.c_float(thirdAlpha),
        C.c_float(fourthAlpha), C.c_float(loadTimeScale), _f4(defoOffset), _f4(defoScale), _f4(defoFactor), _f4(overrideSize),
        C.c_float(overrideTimeOff), int(bordSkip), C.c_float(defoAniFac))
    return out


def load_advect_time_slice_unopt(defo, d3, phi, time, blendAlpha, loadTimeScale, defoOffset, defoScale, defoFactor,
                                 overrideSize=-1., overrideTimeOff=0., defoAniFac=1., zeroVel=False, dst=None):
    """The unoptimised loadAdvectTimeSlice (ref optflow4d.cpp:1671-1760): returns (dst, debugVel (Vec3), debugVelT)."""
    defo = _f32(defo)
    p = _f32(phi)
    out = np.zeros((d3[2], d3[1], d3[0]), np.float32) if dst is None else _f32(dst).copy()
    dv = np.zeros((d3[2], d3[1], d3[0], 3), np.float32)
    dt = np.zeros((d3[2], d3[1], d3[0]), np.float32)
    lib().orc_load_advect_time_slice_unopt(
        _p(defo), _d(dims_of(defo)), _p(out), _p(dv), _p(dt), int(d3[0]), int(d3[1]), int(d3[2]), _p(p), _d(dims_of(p)),
        C.c_float(time), C.c_float(blendAlpha), C.c_float(loadTimeScale), _f4(defoOffset), _f4(defoScale), _f4(defoFactor),
        _f4(overrideSize), C.c_float(overrideTimeOff), C.c_float(defoAniFac), int(bool(zeroVel)))
    return out, dv, dt


def dot_seq(a, b, kind=0, diag=0.0):
    """The reference's sequential dot-product loop (dotProd, optflow4d.cpp:234-241) on Vec4 arrays; kind 1 applies the
    Jacobi preconditioner of grad = b first (precondApply :345-351)."""
    a, b = _f32(a), _f32(b)
    return float(lib().orc_dot_seq(_p(a), _p(b), C.c_longlong(a.size // 4), int(kind), C.c_float(diag)))
