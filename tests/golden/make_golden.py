"""Generates the committed golden fixtures from the REFERENCE ITSELF (oracle/_ref/libofref.so,
i.e. thunil/ofblend compiled unmodified from /root/reference by oracle/Makefile).

Run in the dev container only (the reference does not exist on the GPU box):
    python tests/golden/make_golden.py small          # per-operator arrays, 14x12x13x16
    python tests/golden/make_golden.py mode1 32 48    # synthetic two-drop pair, 32^3 x 48
    python tests/golden/make_golden.py mode1 64 64    # synthetic two-drop pair, 64^4

Fixtures
  small_ops.npz          inputs are re-created from seeds by the tests; outputs stored in full
  mode1_<nx>x<nt>.npz    reference trace of opticalFlowMultiscale4d with the README parameters
                         (CG iterations per solve, error values) + the deformation sub-sampled on
                         a fixed lattice (every `stride` cells) + its L2 norm / max, so that tests
                         at BASELINE.json's sizes need no 268 MB golden file.
"""
import os
import re
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref  # noqa: E402
from ofblend_b200 import synth  # noqa: E402


class RefOps:
    set_bound4d = staticmethod(lambda a, v, w: ref.set_bound4d(a, v, w))
    extrap4d_ls_simple = staticmethod(lambda a, d, i: ref.extrap4d_ls_simple(a, d, i))
    mult_const = staticmethod(lambda a, s: ref.grid_op4d("multConst", a, None, s))


def capture_stdout(fn):
    """Run fn() while capturing the C++ std::cout of the reference (debMsg)."""
    sys.stdout.flush()
    with tempfile.TemporaryFile(mode="w+b") as tf:
        old = os.dup(1)
        os.dup2(tf.fileno(), 1)
        try:
            out = fn()
        finally:
            os.dup2(old, 1)
            os.close(old)
        tf.seek(0)
        return out, tf.read().decode(errors="replace")


def rnd(shape, seed, scale=1.0):
    return (np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32)


def small():
    from conftest import sdf_pair
    D = (14, 12, 13, 16)
    SH = (D[3], D[2], D[1], D[0])
    i0, i1 = sdf_pair(D)
    vel = rnd(SH + (4,), 3, 2.5)
    out = {}
    out["interp_real"] = ref.interpolate_grid4d(i0, (7, 6, 6, 8), (1.4, 1.2, 1.3, 3.0), 0.8)
    out["interp_vec_down"] = ref.interpol_grid_templ(vel, (7, 6, 6, 8))
    out["interp_vec_up"] = ref.interpol_grid_templ(out["interp_vec_down"], D)
    out["advect_real"] = ref.advect4d(vel, i0, 0.7)
    out["advect_vec4"] = ref.advect4d(vel, rnd(SH + (4,), 4), 0.7)
    out["advect_cfl"] = ref.advect_cfl4d(1.5, vel, i0, 1.0)
    out["blur_s2"] = ref.gaussian_blur4d(rnd(SH + (4,), 9), 2.0)
    out["blur_s1"] = ref.gaussian_blur4d(rnd(SH + (4,), 9), 1.125)
    v0 = np.zeros(SH + (4,), np.float32)
    (of, rhs), log = capture_stdout(lambda: (ref.set_debug_level(1), ref.optical_flow4d(
        v0, i0, i1, 1e-3, 1e-4, 0., 1e-2, -1., want_rhs=True), ref.set_debug_level(0))[1])
    out["of_vel"] = of
    out["of_rhs"] = rhs
    out["of_iters"] = np.array([int(x) for x in re.findall(r"ofSolve fix iterations:(\d+)", log)], np.int32)
    velp = rnd(SH + (4,), 11, 0.7)
    out["proj_dst"], out["proj_marker"] = ref.project_cells(velp, i0, i1, 4., 40)
    out["expol5"] = ref.cv_expol_blur4d(out["proj_dst"], out["proj_marker"], 5)
    cd, cv = ref.corr_vels_of4d(np.zeros_like(velp), rnd(SH + (4,), 12, 0.5), i0, i1, 4., 4., 0.1, 40)
    out["corr_dst"], out["corr_vel"] = cd, cv
    out["lsdiff"] = np.array([ref.calc_ls_diff4d(i0, i1, 0.005, 0), ref.calc_ls_diff4d(i0, i1, 0.005, 2)], np.float32)
    phi = ref.set_bound4d(i0 / np.float32(-0.005), 0.1, 1)
    for inside in (0, 1):
        p, m = ref.extrap4d_ls_simple(phi, 6, bool(inside), want_marker=True)
        out["extrap_phi_%d" % inside] = p
        out["extrap_marker_%d" % inside] = m
    out["extrap_vec4"] = ref.extrapolate_vec4_simple(rnd(SH + (4,), 13), phi, 5)
    out["repeat"] = ref.repeat_frame4d(rnd(SH, 14), 4.3, 3.0, 0)
    out["neumann_w1"] = ref.set_bound_neumann4d(vel, 1)
    out["setbound_w3"] = ref.set_bound4d(i0, 0.1, 3)
    a3 = rnd((13, 12, 14), 15)
    out["blur_special"] = ref.simple_blur_special(a3, 2, -999., 1)
    out["checker"] = ref.init_test_checkerboard((8, 8, 8, 8))
    # test_0032_grid4dop.py expected values (analytic): Real 1.1, 1.2, 2.9 ...
    np.savez_compressed(os.path.join(HERE, "small_ops.npz"), **out)
    print("wrote small_ops.npz", {k: v.shape for k, v in out.items()})


def mode1(nx, nt, threads=8):
    ref.set_threads(threads)
    dims = (nx, nx, nx, nt)
    i0 = synth.post_process(synth.two_drop_phi(dims, 0), RefOps)
    i1 = synth.post_process(synth.two_drop_phi(dims, 1), RefOps)
    v0 = np.zeros(i0.shape + (4,), np.float32)
    ref.set_debug_level(1)
    t0 = time.time()
    vel, log = capture_stdout(lambda: ref.optical_flow_multiscale4d(v0, i0, i1, **synth.MODE1_PARAMS))
    wall = time.time() - t0
    ref.set_debug_level(0)
    iters = [int(x) for x in re.findall(r"ofSolve fix iterations:(\d+)", log)]
    cgsec = [float(x) for x in re.findall(r"ofSolve fix iterations:\d+ \(([0-9.eE+-]+)s\)", log)]
    errs = [float(x) for x in re.findall(r"Current error, s\d+ \d+ = ([0-9.eE+-]+)", log)]
    errs += [float(x) for x in re.findall(r"Final error=([0-9.eE+-]+)", log)]
    expol = [int(x) for x in re.findall(r"Extrapol distance increased to (\d+)", log)]
    stride = max(1, nx // 16)
    sub = vel[::stride, ::stride, ::stride, ::stride].copy()
    adv = ref.advect4d(vel, i0)
    # the same run without the final SDF projection: the projection amplifies round-off level
    # differences of its input by ~4e4 (DESIGN.md "conditioning"), so the pre-projection field is
    # stored too and held to the 1e-4 bar on its own
    pnp = dict(synth.MODE1_PARAMS)
    pnp["doFinalProject"] = False
    vel_np = ref.optical_flow_multiscale4d(v0, i0, i1, **pnp)
    adv_np = ref.advect4d(vel_np, i0)
    fn = os.path.join(HERE, "mode1_%dx%d.npz" % (nx, nt))
    np.savez_compressed(fn, dims=np.array(dims, np.int32), cg_iters=np.array(iters, np.int32),
                        cg_seconds=np.array(cgsec, np.float32), errs=np.array(errs, np.float32),
                        expol=np.array(expol, np.int32), stride=np.int32(stride), vel_sub=sub,
                        vel_l2=np.float64(np.linalg.norm(vel.astype(np.float64).ravel())),
                        vel_maxabs=np.float32(np.abs(vel).max()),
                        adv_sub=adv[::stride, ::stride, ::stride, ::stride].copy(),
                        vel_noproj_sub=vel_np[::stride, ::stride, ::stride, ::stride].copy(),
                        vel_noproj_l2=np.float64(np.linalg.norm(vel_np.astype(np.float64).ravel())),
                        adv_noproj_sub=adv_np[::stride, ::stride, ::stride, ::stride].copy(),
                        i0_sum=np.float64(i0.astype(np.float64).sum()), i1_sum=np.float64(i1.astype(np.float64).sum()),
                        ref_wall_s=np.float32(wall), ref_threads=np.int32(threads))
    print("wrote", fn, "iters", iters, "errs", errs, "wall %.1fs" % wall, "threads", threads)
    full = os.environ.get("GOLDEN_FULL_DIR")
    if full:
        np.save(os.path.join(full, "mode1_%dx%d_vel.npy" % (nx, nt)), vel)


if __name__ == "__main__":
    if sys.argv[1] == "small":
        small()
    else:
        mode1(int(sys.argv[2]), int(sys.argv[3]))
