"""ctypes binding of the C ABI (include/flof_b200.h, ofblend_b200/libflof_b200.so).

This is the *Python-side* stub a maintainer would use to reach the B200 kernels (the C++ host
layer `manta` module binds the same symbols); tests/ and bench.py drive the product through it.
It never imports anything from oracle/ and has no CPU fallback: if the shared library or a CUDA
device is missing, Context() raises.

numpy layout: array[t, z, y, x(, 4)], byte-identical to the reference's Grid4d<T>.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLOF_B200_LIB") or os.path.join(_HERE, "libflof_b200.so")  # env override: A/B builds


class Dim4(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("nt", C.c_int)]


class Dim3(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int)]


class MultiscaleParams(C.Structure):
    _fields_ = [("wSmooth", C.c_float), ("wEnergy", C.c_float), ("postVelBlur", C.c_float),
                ("cgAccuracy", C.c_float), ("cfl", C.c_float), ("resetBndWidth", C.c_float),
                ("multiStep", C.c_int), ("projSizeThresh", C.c_int), ("minGridSize", C.c_int),
                ("doFinalProject", C.c_int)]


class MultiscaleTrace(C.Structure):
    _fields_ = [("n_solves", C.c_int), ("cg_iters", C.c_int * 64), ("cg_ms", C.c_float * 64),
                ("cg_cells", C.c_int64 * 64), ("n_errs", C.c_int), ("errs", C.c_float * 64),
                ("total_ms", C.c_float)]


class FlofError(RuntimeError):
    pass


_lib = None


def load_library():
    """dlopen libflof_b200.so (no device needed); raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise FlofError("%s missing: build it with `make -C ofblend_b200/csrc` "
                            "(__graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.flof_last_error.restype = C.c_char_p
        _lib.flof_last_error.argtypes = [C.c_void_p]
        _lib.flof_ctx_stream.restype = C.c_void_p
        _lib.flof_ctx_launch_count.restype = C.c_longlong
    return _lib


def _f4(v):
    v = np.broadcast_to(np.asarray(v, dtype=np.float32), (4,))
    return (C.c_float * 4)(*[float(x) for x in v])


def _f3(v):
    v = np.broadcast_to(np.asarray(v, dtype=np.float32), (3,))
    return (C.c_float * 3)(*[float(x) for x in v])


def dims_of(a):
    s = a.shape[:4]
    return (s[3], s[2], s[1], s[0])


class DeviceGrid:
    """A device-resident Grid4d<T> / Grid<T> buffer owned by a Context."""

    def __init__(self, ctx, dims, elem, dtype=np.float32, ptr=None):
        self.ctx = ctx
        self.dims = tuple(int(x) for x in dims)
        self.elem = int(elem)
        self.dtype = np.dtype(dtype)
        self.cells = int(np.prod(self.dims))
        self.nbytes = self.cells * self.elem * self.dtype.itemsize
        self.ptr = C.c_void_p()
        ctx._chk(ctx.lib.flof_malloc(ctx.h, C.byref(self.ptr), C.c_size_t(self.nbytes)))

    @property
    def shape(self):
        return tuple(reversed(self.dims)) + ((self.elem,) if self.elem > 1 else ())

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        assert a.size * a.itemsize == self.nbytes, (a.shape, self.shape)
        self.ctx._chk(self.ctx.lib.flof_memcpy_h2d(self.ctx.h, self.ptr, a.ctypes.data_as(C.c_void_p),
                                                   C.c_size_t(self.nbytes)))
        self.ctx.sync()  # `a` may be a temporary
        return self

    def download(self):
        out = np.empty(self.shape, self.dtype)
        self.ctx._chk(self.ctx.lib.flof_memcpy_d2h(self.ctx.h, out.ctypes.data_as(C.c_void_p), self.ptr,
                                                   C.c_size_t(self.nbytes)))
        return out

    def free(self):
        if self.ptr:
            self.ctx.lib.flof_free(self.ctx.h, self.ptr)
            self.ptr = C.c_void_p()

    def d4(self):
        return Dim4(*self.dims)

    def d3(self):
        return Dim3(*self.dims[:3])


class Context:
    """flof_ctx: one CUDA device + stream + grid memory pool."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.flof_ctx_create(C.byref(self.h), int(device))
        if rc != 0:
            raise FlofError("flof_ctx_create failed: " + self.lib.flof_last_error(None).decode())

    def close(self):
        if self.h:
            self.lib.flof_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def _chk(self, rc):
        if rc != 0:
            raise FlofError(self.lib.flof_last_error(self.h).decode())

    def sync(self):
        self._chk(self.lib.flof_sync(self.h))

    def set_option(self, name, value):
        """Context option (flof_ctx_set_option).  Kernel choices that are bit-identical: expol_mode, expol_variant, apply_variant,
        apply_zchunk, sweep_overlap; multi-GPU plumbing: no_p2p, host_result_rank; opt-in arithmetic changes (NOT the reference's
        bits): dot_mode 0 (tree-reduced CG dot products), blur_mode 1 (separable Gaussian)."""
        self._chk(self.lib.flof_ctx_set_option(self.h, name.encode(), int(value)))

    def get_option(self, name):
        v = C.c_int(0)
        self._chk(self.lib.flof_ctx_get_option(self.h, name.encode(), C.byref(v)))
        return int(v.value)

    def dot_seq(self, a, b, kind=0, diag=0.0):
        """CG dot product in the reference's sequential summation order (flof_dot_seq); a, b: device Vec4 grids."""
        out = C.c_double(0)
        st = (C.c_ulonglong * 15)()
        self._chk(self.lib.flof_dot_seq(self.h, a.ptr, b.ptr, C.c_int64(a.cells), int(kind), C.c_float(diag), C.byref(out), st))
        return out.value, [int(x) for x in st]

    def seq_stats(self):
        st = (C.c_ulonglong * 15)()
        self._chk(self.lib.flof_seq_stats(self.h, st))
        return dict(zip(("dots", "dirty_leaves", "raw_products", "pieces", "fallbacks", "inconsistent", "careful_segments", "inexact", "why", "raw_leaves",
                         "cyc_gather", "cyc_compose", "cyc_walk", "cyc_finish", "walk_steps"),
                        [int(x) for x in st]))

    @property
    def stream(self):
        return self.lib.flof_ctx_stream(self.h)

    @property
    def sm_count(self):
        return int(self.lib.flof_ctx_sm_count(self.h))

    @property
    def launches(self):
        return int(self.lib.flof_ctx_launch_count(self.h))

    # ---- allocation helpers
    def grid(self, dims, elem=1, dtype=np.float32):
        return DeviceGrid(self, dims, elem, dtype)

    def to_device(self, a):
        a = np.asarray(a)
        if a.dtype == np.int32:
            return DeviceGrid(self, dims_of(a) if a.ndim == 4 else tuple(reversed(a.shape)), 1, np.int32).upload(a)
        a = np.ascontiguousarray(a, np.float32)
        if a.ndim == 5:
            return DeviceGrid(self, dims_of(a), a.shape[4]).upload(a)
        if a.ndim == 4:
            return DeviceGrid(self, dims_of(a), 1).upload(a)
        if a.ndim == 3:
            return DeviceGrid(self, (a.shape[2], a.shape[1], a.shape[0]), 1).upload(a)
        raise ValueError("unsupported array rank %d" % a.ndim)

    # ---- thin wrappers over the C ABI (device grids in, device grids out)
    def grid_binary(self, a, b, op):
        self._chk(self.lib.flof_grid_binary(self.h, a.ptr, b.ptr, C.c_int64(a.cells), a.elem, int(op)))

    def grid_add_scaled(self, a, b, f):
        self._chk(self.lib.flof_grid_add_scaled(self.h, a.ptr, b.ptr, C.c_int64(a.cells), a.elem, _f4(f)))

    def grid_mult_const(self, a, f):
        self._chk(self.lib.flof_grid_mult_const(self.h, a.ptr, C.c_int64(a.cells), a.elem, _f4(f)))

    def grid_add_const(self, a, f):
        self._chk(self.lib.flof_grid_add_const(self.h, a.ptr, C.c_int64(a.cells), a.elem, _f4(f)))

    def grid_set_const(self, a, f):
        self._chk(self.lib.flof_grid_set_const(self.h, a.ptr, C.c_int64(a.cells), a.elem, _f4(f)))

    def grid_clamp(self, a, lo, hi):
        self._chk(self.lib.flof_grid_clamp(self.h, a.ptr, C.c_int64(a.cells), a.elem, C.c_float(lo), C.c_float(hi)))

    def grid_min_max(self, a):
        out = (C.c_float * 3)()
        self._chk(self.lib.flof_grid_min_max(self.h, a.ptr, C.c_int64(a.cells), a.elem, out))
        return tuple(float(x) for x in out)

    def grid_max_diff(self, a, b):
        out = C.c_double(0)
        self._chk(self.lib.flof_grid_max_diff(self.h, a.ptr, b.ptr, C.c_int64(a.cells), a.elem, C.byref(out)))
        return out.value

    def set_bound(self, a, v, w=1):
        if a.dtype == np.int32:
            self._chk(self.lib.flof_grid4d_set_bound_int(self.h, a.ptr, a.d4(), int(np.atleast_1d(v)[0]), int(w)))
        else:
            self._chk(self.lib.flof_grid4d_set_bound(self.h, a.ptr, a.d4(), a.elem, _f4(v), int(w)))

    def set_bound_neumann(self, a, w=1):
        self._chk(self.lib.flof_grid4d_set_bound_neumann(self.h, a.ptr, a.d4(), a.elem, int(w)))

    def interpolate_grid4d(self, dst, src, offset=0., scale=1., size=-1.):
        self._chk(self.lib.flof_interpolate_grid4d(self.h, dst.ptr, dst.d4(), src.ptr, src.d4(), src.elem,
                                                   _f4(offset), _f4(scale), _f4(size)))

    def interpol_grid_templ(self, dst, src):
        self._chk(self.lib.flof_interpol_grid_templ(self.h, dst.ptr, dst.d4(), src.ptr, src.d4(), src.elem))

    def semi_lagrange4d(self, vel, src, dst, dt):
        self._chk(self.lib.flof_semi_lagrange4d(self.h, vel.ptr, src.ptr, dst.ptr, src.d4(), src.elem, C.c_float(dt)))

    def advect4d(self, vel, grid, dt):
        self._chk(self.lib.flof_advect4d(self.h, vel.ptr, grid.ptr, grid.d4(), grid.elem, C.c_float(dt)))

    def advect_cfl4d(self, cfl, vel, grid, velFactor=1.):
        self._chk(self.lib.flof_advect_cfl4d(self.h, C.c_float(cfl), vel.ptr, grid.ptr, grid.d4(), grid.elem,
                                             C.c_float(velFactor)))

    def of_assemble(self, grad, rhs, i0, i1, vel, wSmooth, wEnergy):
        self._chk(self.lib.flof_of_assemble(self.h, grad.ptr, rhs.ptr, i0.ptr, i1.ptr,
                                            vel.ptr if vel is not None else None, i0.d4(),
                                            C.c_float(wSmooth), C.c_float(wEnergy)))

    def of_cg(self, x, grad, rhs, wSmooth, wEnergy, accuracy, maxIter=1000):
        it = C.c_int(0)
        rr = C.c_float(0)
        self._chk(self.lib.flof_of_cg(self.h, x.ptr, grad.ptr, rhs.ptr, x.d4(), C.c_float(wSmooth),
                                      C.c_float(wEnergy), C.c_float(accuracy), int(maxIter), C.byref(it),
                                      C.byref(rr)))
        return it.value, rr.value

    def optical_flow4d(self, vel, i0, i1, rhsT=None, wSmooth=0., wEnergy=0., postVelBlur=0., cgAccuracy=1e-4,
                       resetBndWidth=-1.):
        it = C.c_int(0)
        rr = C.c_float(0)
        self._chk(self.lib.flof_optical_flow4d(self.h, vel.ptr, i0.ptr, i1.ptr, rhsT.ptr if rhsT else None,
                                               i0.d4(), C.c_float(wSmooth), C.c_float(wEnergy),
                                               C.c_float(postVelBlur), C.c_float(cgAccuracy),
                                               C.c_float(resetBndWidth), C.byref(it), C.byref(rr)))
        return it.value, rr.value

    def gaussian_blur4d(self, a, sigma, iters=1):
        self._chk(self.lib.flof_gaussian_blur4d(self.h, a.ptr, a.d4(), a.elem, C.c_float(sigma), int(iters)))

    def cv_expol_blur4d(self, a, marker, sweeps):
        self._chk(self.lib.flof_cv_expol_blur4d(self.h, a.ptr, marker.ptr, a.d4(), int(sweeps)))

    def project_cells(self, dst, vel, phiOrg, phiTarget, marker, threshPhi, maxIter):
        self._chk(self.lib.flof_project_cells(self.h, dst.ptr, vel.ptr, phiOrg.ptr, phiTarget.ptr, marker.ptr,
                                              vel.d4(), C.c_float(threshPhi), int(maxIter)))

    def corr_vels_of4d(self, dst, vel, phiOrg, phiTarget, threshPhi=1e10, postVelBlur=0., resetBndWidth=-1.,
                       maxIter=100):
        self._chk(self.lib.flof_corr_vels_of4d(self.h, dst.ptr, vel.ptr, phiOrg.ptr, phiTarget.ptr, vel.d4(),
                                               C.c_float(threshPhi), C.c_float(postVelBlur),
                                               C.c_float(resetBndWidth), int(maxIter)))

    def calc_ls_diff4d(self, i0, i1, out=None, correction=1., bnd=0):
        r = C.c_float(0)
        self._chk(self.lib.flof_calc_ls_diff4d(self.h, i0.ptr, i1.ptr, out.ptr if out else None, i0.d4(),
                                               C.c_float(correction), int(bnd), C.byref(r)))
        return r.value

    def calc_smoke_diff4d(self, i0, i1, correction=1., bnd=0):
        r = C.c_float(0)
        self._chk(self.lib.flof_calc_smoke_diff4d(self.h, i0.ptr, i1.ptr, i0.d4(), C.c_float(correction), int(bnd), C.byref(r)))
        return r.value

    def optical_flow_multiscale4d(self, vel, i0, i1, params, want_trace=False):
        tr = MultiscaleTrace()
        err = C.c_float(0)
        self._chk(self.lib.flof_optical_flow_multiscale4d(self.h, vel.ptr, i0.ptr, i1.ptr, i0.d4(),
                                                          C.byref(params), C.byref(tr), C.byref(err)))
        return (err.value, tr) if want_trace else err.value

    def optical_flow_multiscale4d_host(self, vel_h, i0_h, i1_h, params):
        """The end-to-end plugin call: HOST buffers in, HOST deformation out (vel_h updated in place)."""
        tr = MultiscaleTrace()
        err = C.c_float(0)
        d = Dim4(*dims_of(i0_h))
        self._chk(self.lib.flof_optical_flow_multiscale4d_host(
            self.h, vel_h.ctypes.data_as(C.c_void_p), i0_h.ctypes.data_as(C.c_void_p),
            i1_h.ctypes.data_as(C.c_void_p), d, C.byref(params), C.byref(tr), C.byref(err)))
        return err.value, tr

    # ---- 3D instantiations: grids created with ctx.grid3(dims3, elem), velocities elem = 3 (Vec3 AoS)
    def optical_flow_multiscale3d(self, vel, i0, i1, params, want_trace=False):
        tr = MultiscaleTrace()
        err = C.c_float(0)
        self._chk(self.lib.flof_optical_flow_multiscale3d(self.h, vel.ptr, i0.ptr, i1.ptr, i0.d3(), C.byref(params),
                                                          C.byref(tr), C.byref(err)))
        return (err.value, tr) if want_trace else err.value

    def corr_vels_of3d(self, dst, vel, phi_org, phi_target, thresh_phi=1e10, post_vel_blur=0., reset_bnd_width=-1.,
                       max_iter=100):
        self._chk(self.lib.flof_corr_vels_of3d(self.h, dst.ptr, vel.ptr, phi_org.ptr, phi_target.ptr, phi_org.d3(),
                                               C.c_float(thresh_phi), C.c_float(post_vel_blur), C.c_float(reset_bnd_width),
                                               int(max_iter)))

    def advect_semi_lagrange_cfl3d(self, cfl, vel, grid, vel_factor=1.):
        self._chk(self.lib.flof_advect_semi_lagrange_cfl3d(self.h, C.c_float(cfl), vel.ptr, grid.ptr, int(grid.elem),
                                                           vel.d3(), C.c_float(vel_factor)))

    def calc_ls_diff3d(self, i0, i1, out=None, correction=1., bnd=0):
        r = C.c_float(0)
        self._chk(self.lib.flof_calc_ls_diff3d(self.h, i0.ptr, i1.ptr, out.ptr if out else None, i0.d3(),
                                               C.c_float(correction), int(bnd), C.byref(r)))
        return r.value

    def repeat_frame4d(self, phi, srct, rng=0., bnd=0):
        self._chk(self.lib.flof_repeat_frame4d(self.h, phi.ptr, phi.d4(), C.c_float(srct), C.c_float(rng), int(bnd)))

    def extrap4d_ls_simple(self, phi, distance=4, inside=False, marker=None):
        self._chk(self.lib.flof_extrap4d_ls_simple(self.h, phi.ptr, phi.d4(), int(distance), int(bool(inside)),
                                                   marker.ptr if marker else None))

    def extrapolate_vec4_simple(self, vel, phi, distance):
        self._chk(self.lib.flof_extrapolate_vec4_simple(self.h, vel.ptr, phi.ptr, phi.d4(), int(distance)))

    def simple_blur_special(self, a, iters=1, thresh=0., bord=0):
        self._chk(self.lib.flof_simple_blur_special(self.h, a.ptr, a.d3(), int(iters), C.c_float(thresh), int(bord)))

    def grid3_set_bound(self, a, v, w=1):
        self._chk(self.lib.flof_grid3_set_bound(self.h, a.ptr, a.d3(), C.c_float(v), int(w)))


def make_params(wSmooth=0., wEnergy=0., postVelBlur=0., cgAccuracy=1e-4, cfl=999., resetBndWidth=-1., multiStep=1,
                projSizeThresh=9999, minGridSize=10, doFinalProject=False):
    """kwargs of opticalFlowMultiscale4d (ref optflow4d.cpp:2182-2189) -> flof_multiscale_params"""
    return MultiscaleParams(wSmooth, wEnergy, postVelBlur, cgAccuracy, cfl, resetBndWidth, int(multiStep),
                            int(projSizeThresh), int(minGridSize), int(bool(doFinalProject)))


class HostAPI:
    """numpy-in / numpy-out mirror of the reference plugin functions, same surface as oracle/ref.py,
    so parity tests read `gpu.advect4d(...) == oracle.advect4d(...)`.  Every call uploads its
    arguments, runs the CUDA path through the C ABI and downloads the result."""

    def __init__(self, ctx=None, device=0):
        self.ctx = ctx or Context(device)

    def _run(self, arrays, fn):
        ctx = self.ctx
        devs = [ctx.to_device(a) if a is not None else None for a in arrays]
        try:
            return fn(*devs)
        finally:
            for g in devs:
                if g is not None:
                    g.free()

    def interpolate_grid4d(self, src, tdims, offset=0., scale=1., size=-1.):
        def f(s):
            dst = self.ctx.grid(tdims, s.elem)
            self.ctx.interpolate_grid4d(dst, s, offset, scale, size)
            out = dst.download()
            dst.free()
            return out
        return self._run([src], f)

    def interpol_grid_templ(self, src, tdims):
        def f(s):
            dst = self.ctx.grid(tdims, s.elem)
            self.ctx.interpol_grid_templ(dst, s)
            out = dst.download()
            dst.free()
            return out
        return self._run([src], f)

    def advect4d(self, vel, grid, dtFac=1., dt=1.):
        def f(v, g):
            self.ctx.advect4d(v, g, float(np.float32(dt) * np.float32(dtFac)))
            return g.download()
        return self._run([vel, grid], f)

    def advect_cfl4d(self, cfl, vel, grid, velFactor=1.):
        def f(v, g):
            self.ctx.advect_cfl4d(cfl, v, g, velFactor)
            return g.download()
        return self._run([vel, grid], f)

    def optical_flow4d(self, vel, i0, i1, wSmooth=0., wEnergy=0., postVelBlur=0., cgAccuracy=1e-4,
                       resetBndWidth=-1., want_rhs=False, want_iters=False):
        def f(v, a, b):
            rhs = self.ctx.grid(a.dims, 1) if want_rhs else None
            it, rr = self.ctx.optical_flow4d(v, a, b, rhs, wSmooth, wEnergy, postVelBlur, cgAccuracy, resetBndWidth)
            out = (v.download(),)
            if want_rhs:
                out += (rhs.download(),)
                rhs.free()
            if want_iters:
                out += (it,)
            return out if len(out) > 1 else out[0]
        return self._run([vel, i0, i1], f)

    def gaussian_blur4d(self, a, sigma, iters=1):
        def f(g):
            self.ctx.gaussian_blur4d(g, sigma, iters)
            return g.download()
        return self._run([a], f)

    def project_cells(self, vel, phiOrg, phiTarget, threshPhi, maxIter):
        def f(v, po, pt):
            dst = self.ctx.grid(v.dims, 4)
            mk = self.ctx.grid(v.dims, 1)
            self.ctx.project_cells(dst, v, po, pt, mk, threshPhi, maxIter)
            out = (dst.download(), mk.download())
            dst.free()
            mk.free()
            return out
        return self._run([vel, phiOrg, phiTarget], f)

    def cv_expol_blur4d(self, a, marker, sweeps):
        def f(g, m):
            self.ctx.cv_expol_blur4d(g, m, sweeps)
            return g.download()
        return self._run([a, marker], f)

    def corr_vels_of4d(self, dst, vel, phiOrg, phiTarget, threshPhi=1e10, postVelBlur=0., resetBndWidth=-1.,
                       maxIter=100):
        def f(d, v, po, pt):
            self.ctx.corr_vels_of4d(d, v, po, pt, threshPhi, postVelBlur, resetBndWidth, maxIter)
            return d.download(), v.download()
        return self._run([dst, vel, phiOrg, phiTarget], f)

    def calc_ls_diff4d(self, i0, i1, correction=1., bnd=0, want_out=False):
        def f(a, b):
            out = self.ctx.grid(a.dims, 1) if want_out else None
            r = self.ctx.calc_ls_diff4d(a, b, out, correction, bnd)
            if want_out:
                o = out.download()
                out.free()
                return r, o
            return r
        return self._run([i0, i1], f)

    def calc_smoke_diff4d(self, i0, i1, correction=1., bnd=0):
        return self._run([i0, i1], lambda a, b: self.ctx.calc_smoke_diff4d(a, b, correction, bnd))

    def optical_flow_multiscale4d(self, vel, i0, i1, want_trace=False, **kw):
        p = make_params(**kw)

        def f(v, a, b):
            err, tr = self.ctx.optical_flow_multiscale4d(v, a, b, p, want_trace=True)
            out = v.download()
            if want_trace:
                return out, list(tr.cg_iters[:tr.n_solves]), [float(x) for x in tr.errs[:tr.n_errs]]
            return out
        return self._run([vel, i0, i1], f)

    # ---- 3D instantiations: numpy layout [z, y, x(, 3)]
    def _g3(self, a):
        a = np.ascontiguousarray(a, np.float32)
        return DeviceGrid(self.ctx, (a.shape[2], a.shape[1], a.shape[0]), 3 if a.ndim == 4 else 1).upload(a)

    def optical_flow_multiscale3d(self, vel, i0, i1, want_trace=False, **kw):
        p = make_params(**kw)
        gs = [self._g3(vel), self._g3(i0), self._g3(i1)]
        try:
            err, tr = self.ctx.optical_flow_multiscale3d(gs[0], gs[1], gs[2], p, want_trace=True)
            out = gs[0].download()
            if want_trace:
                return out, list(tr.cg_iters[:tr.n_solves]), [float(x) for x in tr.errs[:tr.n_errs]]
            return out
        finally:
            for g in gs:
                g.free()

    def corr_vels_of3d(self, dst, vel, phiOrg, phiTarget, threshPhi=1e10, postVelBlur=0., resetBndWidth=-1., maxIter=100):
        gs = [self._g3(dst), self._g3(vel), self._g3(phiOrg), self._g3(phiTarget)]
        try:
            self.ctx.corr_vels_of3d(gs[0], gs[1], gs[2], gs[3], threshPhi, postVelBlur, resetBndWidth, maxIter)
            return gs[0].download(), gs[1].download()
        finally:
            for g in gs:
                g.free()

    def advect_semi_lagrange_cfl3d(self, cfl, vel, grid, velFactor=1.):
        gs = [self._g3(vel), self._g3(grid)]
        try:
            self.ctx.advect_semi_lagrange_cfl3d(cfl, gs[0], gs[1], velFactor)
            return gs[1].download()
        finally:
            for g in gs:
                g.free()

    def calc_ls_diff3d(self, i0, i1, correction=1., bnd=0, want_out=False):
        gs = [self._g3(i0), self._g3(i1)]
        out = DeviceGrid(self.ctx, gs[0].dims, 1) if want_out else None
        try:
            r = self.ctx.calc_ls_diff3d(gs[0], gs[1], out, correction, bnd)
            return (r, out.download()) if want_out else r
        finally:
            for g in gs + ([out] if out else []):
                g.free()

    def extrap4d_ls_simple(self, phi, distance=4, inside=False, want_marker=False):
        def f(p):
            mk = self.ctx.grid(p.dims, 1, np.int32) if want_marker else None
            self.ctx.extrap4d_ls_simple(p, distance, inside, mk)
            if want_marker:
                out = (p.download(), mk.download())
                mk.free()
                return out
            return p.download()
        return self._run([phi], f)

    def extrapolate_vec4_simple(self, vel, phi, distance):
        def f(v, p):
            self.ctx.extrapolate_vec4_simple(v, p, distance)
            return v.download()
        return self._run([vel, phi], f)

    def repeat_frame4d(self, phi, srct, rng=0., bnd=0):
        def f(p):
            self.ctx.repeat_frame4d(p, srct, rng, bnd)
            return p.download()
        return self._run([phi], f)

    def set_bound4d(self, a, value, w=1):
        def f(g):
            self.ctx.set_bound(g, value, w)
            return g.download()
        return self._run([a], f)

    def set_bound_neumann4d(self, a, w=1):
        def f(g):
            self.ctx.set_bound_neumann(g, w)
            return g.download()
        return self._run([a], f)

    def min_max4d(self, a):
        return self._run([a], lambda g: self.ctx.grid_min_max(g))

    def grid_op4d(self, op, a, b=None, factor=0.):
        def f(ga, gb):
            c = self.ctx
            if op == "add":
                c.grid_binary(ga, gb, 0)
            elif op == "sub":
                c.grid_binary(ga, gb, 1)
            elif op == "mult":
                c.grid_binary(ga, gb, 2)
            elif op == "addScaled":
                c.grid_add_scaled(ga, gb, factor)
            elif op == "multConst":
                c.grid_mult_const(ga, factor)
            elif op == "addConst":
                c.grid_add_const(ga, factor)
            elif op == "clamp":
                c.grid_clamp(ga, factor[0], factor[1])
            else:
                raise ValueError(op)
            return ga.download()
        return self._run([a, b], f)

    def mult_const(self, a, s):
        return self.grid_op4d("multConst", a, None, s)

    def simple_blur_special(self, a, iters=1, thresh=0., bord=0):
        def f(g):
            self.ctx.simple_blur_special(g, iters, thresh, bord)
            return g.download()
        return self._run([a], f)

    def grid3_set_bound(self, a, value, w=1):
        def f(g):
            self.ctx.grid3_set_bound(g, value, w)
            return g.download()
        return self._run([a], f)

    def levelset_join(self, a, b):
        def f(ga, gb):
            self.ctx.grid_binary(ga, gb, 3)
            return ga.download()
        return self._run([a, b], f)

    def load_place_grid4d(self, slices, phi, offset, scale, fileIdxStart=-1, fileIdxEnd=-1, debugSkipLoad=999999,
                          spread=1., overrideSize=-1., overrideTimeOff=0., overrideGoodRegion=0, loadTimeScale=1.,
                          rescaleSdfValues=False, sdfIsoOff=0., repeatStartFrame=0.):
        slices = np.ascontiguousarray(slices, np.float32)
        ctx = self.ctx
        sl = DeviceGrid(ctx, (slices.shape[3], slices.shape[2], slices.shape[1], slices.shape[0]), 1).upload(slices)
        p = ctx.to_device(phi)
        try:
            ctx._chk(ctx.lib.flof_load_place_grid4d(
                ctx.h, sl.ptr, int(slices.shape[0]), Dim3(slices.shape[3], slices.shape[2], slices.shape[1]), p.ptr,
                p.d4(), _f4(offset), _f4(scale), int(fileIdxStart), int(fileIdxEnd), int(debugSkipLoad),
                C.c_float(spread), _f4(overrideSize), C.c_float(overrideTimeOff), int(overrideGoodRegion),
                C.c_float(loadTimeScale), int(bool(rescaleSdfValues)), C.c_float(sdfIsoOff),
                C.c_float(repeatStartFrame)))
            return p.download()
        finally:
            sl.free()
            p.free()

    def shift_forw_grid4d(self, phi, overrideGoodRegion):
        def f(p):
            self.ctx._chk(self.ctx.lib.flof_shift_forw_grid4d(self.ctx.h, p.ptr, p.d4(), int(overrideGoodRegion)))
            return p.download()
        return self._run([phi], f)

    def load_advect_defovols(self, vols, d3, phi, times, blendAlpha, thirdAlpha, fourthAlpha, loadTimeScale, defoOffset,
                             defoScale, defoFactor, doAligned=False, partialLoadFac=0.2, overrideSize=-1., overrideTimeOff=0.,
                             bordSkip=1, defoAniFac=1.):
        """n frames through 2 / 3 deformation volumes, orchestrated like the `manta` module's _OptInit(useDefoVols=True) /
        _OptAdd / _OptRun: window refresh, composition of the one-slice field, look-up (ref optflow4d.cpp:1822-1863,
        1951-2105)."""
        ctx = self.ctx
        lib = ctx.lib

        def f(p, *dv):
            dd = dv[0].d4()
            dimT = int(dd.nt)
            Tw = int(np.float32(dimT) * max(np.float32(0.2), np.float32(partialLoadFac)))
            wdims = (int(dd.nx), int(dd.ny), int(dd.nz), Tw)
            wins = [ctx.grid(wdims, 4) for _ in dv]
            for w in wins:
                ctx._chk(lib.flof_grid_set_const(ctx.h, w.ptr, C.c_int64(w.cells), 4, _f4(0.)))
            dvt = ctx.grid(wdims, 4) if doAligned else None
            vt = ctx.grid(wdims[:3] + (1,), 4)
            out = ctx.grid((d3[0], d3[1], d3[2], 1), 1)
            frames = []
            lastT = -1
            filepos = [C.c_int(-1) for _ in dv]
            ctx._chk(lib.flof_grid_set_const(ctx.h, vt.ptr, C.c_int64(vt.cells), 4, _f4(0.)))
            fac = _f4(np.asarray(np.broadcast_to(np.asarray(defoFactor, np.float32), (4,))) * np.float32(defoAniFac))
            for time in times:
                srcTime, t = C.c_float(0), C.c_int(0)
                sf3, off3 = (C.c_float * 3)(), (C.c_float * 3)()
                lib.flof_lats_source_time(dd, p.d4(), C.c_float(time), C.c_float(loadTimeScale), _f4(defoOffset), _f4(defoScale),
                                          _f4(overrideSize), C.byref(srcTime), C.byref(t), None, None, sf3, off3)
                for w, v, fp in zip(wins, dv, filepos):
                    ctx._chk(lib.flof_defovol_window_update(ctx.h, w.ptr, Dim4(*wdims), v.ptr, dimT, t.value, lastT, vt.ptr, C.byref(fp)))
                lastT = t.value
                tcoord = np.float32(srcTime.value) - np.float32(t.value - Tw // 2)
                ctx._chk(lib.flof_defovol_compose(ctx.h, vt.ptr, wins[0].ptr, wins[1].ptr, wins[2].ptr if len(wins) > 2 else None,
                                                  dvt.ptr if dvt else None, Dim4(*wdims), C.c_float(tcoord), int(bool(doAligned)),
                                                  C.c_float(blendAlpha), C.c_float(thirdAlpha), C.c_float(fourthAlpha)))
                ctx._chk(lib.flof_grid_set_const(ctx.h, out.ptr, C.c_int64(out.cells), 1, _f4(0.)))
                ctx._chk(lib.flof_lookup_slice4d_with_vel(
                    ctx.h, out.ptr, Dim3(*[int(x) for x in d3]), p.ptr, p.d4(), C.c_float(np.float32(time) + np.float32(overrideTimeOff)),
                    C.c_float(1.0), vt.ptr, Dim3(*wdims[:3]), sf3, off3, fac, int(bordSkip)))
                frames.append(out.download().reshape(d3[2], d3[1], d3[0]))
            for g in wins + [vt, out] + ([dvt] if dvt else []):
                g.free()
            return np.stack(frames)
        return self._run([phi] + list(vols), f)

    def load_advect_time_slice_unopt(self, defo, d3, phi, time, blendAlpha, loadTimeScale, defoOffset, defoScale, defoFactor,
                                     overrideSize=-1., overrideTimeOff=0., defoAniFac=1., zeroVel=False):
        """The unoptimised loadAdvectTimeSlice: returns (dst, debugVel, debugVelT)."""
        ctx = self.ctx

        def f(dv, p):
            out = ctx.grid((d3[0], d3[1], d3[2], 1), 1)
            dbg3 = ctx.grid((d3[0] * 3, d3[1], d3[2], 1), 1)     # Vec3 grid: 3 floats per cell
            dbgt = ctx.grid((d3[0], d3[1], d3[2], 1), 1)
            ctx._chk(ctx.lib.flof_grid_set_const(ctx.h, out.ptr, C.c_int64(out.cells), 1, _f4(0.)))
            ctx._chk(ctx.lib.flof_load_advect_time_slice_unopt(
                ctx.h, dv.ptr, dv.d4(), out.ptr, Dim3(*[int(x) for x in d3]), p.ptr, p.d4(), C.c_float(time),
                C.c_float(blendAlpha), C.c_float(loadTimeScale), _f4(defoOffset), _f4(defoScale), _f4(defoFactor),
                _f4(overrideSize), C.c_float(overrideTimeOff), C.c_float(defoAniFac), int(bool(zeroVel)), dbg3.ptr, dbgt.ptr))
            res = (out.download().reshape(d3[2], d3[1], d3[0]), dbg3.download().reshape(d3[2], d3[1], d3[0], 3),
                   dbgt.download().reshape(d3[2], d3[1], d3[0]))
            for g in (out, dbg3, dbgt):
                g.free()
            return res
        return self._run([defo, phi], f)

    def load_advect_time_slice(self, defo, d3, phi, time, blendAlpha, loadTimeScale, defoOffset, defoScale,
                               defoFactor, overrideSize=-1., overrideTimeOff=0., bordSkip=1, defoAniFac=1., dst=None):
        ctx = self.ctx

        def f(dv, p):
            out = ctx.grid((d3[0], d3[1], d3[2], 1), 1)
            if dst is not None:
                out.upload(dst)
            ctx._chk(ctx.lib.flof_load_advect_time_slice(
                ctx.h, dv.ptr, dv.d4(), out.ptr, Dim3(*[int(x) for x in d3]), p.ptr, p.d4(), C.c_float(time),
                C.c_float(blendAlpha), C.c_float(loadTimeScale), _f4(defoOffset), _f4(defoScale), _f4(defoFactor),
                _f4(overrideSize), C.c_float(overrideTimeOff), int(bordSkip), C.c_float(defoAniFac)))
            o = out.download().reshape(d3[2], d3[1], d3[0])
            out.free()
            return o
        return self._run([defo, phi], f)
