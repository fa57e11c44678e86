"""Golden fixtures of the 3D instantiations, generated from the REFERENCE ITSELF (oracle/_ref/libofref.so):
    python tests/golden/make_golden3d.py
writes tests/golden/dim3_ops.npz.  Inputs are re-created from seeds by the tests (tests/conftest.py: sdf_pair3, rnd)."""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

from oracle import ref  # noqa: E402
from make_golden import capture_stdout, rnd  # noqa: E402
from conftest import sdf_pair3, DIM3_CASES  # noqa: E402


def main():
    out = {}
    D = (20, 18, 16)
    SH = (D[2], D[1], D[0])
    i0, i1 = sdf_pair3(D)
    vel = rnd(SH + (3,), 3, 2.0)
    r2, o2 = ref.calc_ls_diff3d(i0, i1, 0.005 * 200, 2, want_out=True)
    out["lsdiff"] = np.array([ref.calc_ls_diff3d(i0, i1, 0.005 * 200, 0), r2], np.float32)
    out["lsdiff_out"] = o2
    out["adv_real"] = ref.advect_semi_lagrange_cfl3d(999., vel, i0)
    out["adv_vec3"] = ref.advect_semi_lagrange_cfl3d(999., vel, rnd(SH + (3,), 4))
    out["adv_real_cfl1"] = ref.advect_semi_lagrange_cfl3d(1.0, vel, i0)
    out["adv_real_fac"] = ref.advect_semi_lagrange_cfl3d(1.5, vel, i0, 0.37)
    cd, cv = ref.corr_vels_of3d(np.zeros(SH + (3,), np.float32), rnd(SH + (3,), 12, 0.5), i0, i1, 4., 2., 0.1, 40)
    out["corr_dst"], out["corr_vel"] = cd, cv
    for name, (dims, params) in DIM3_CASES.items():
        a, b = sdf_pair3(dims)
        v0 = np.zeros(a.shape + (3,), np.float32)
        ref.set_debug_level(1)
        v, log = capture_stdout(lambda: ref.optical_flow_multiscale3d(v0, a, b, **params))
        ref.set_debug_level(0)
        iters = [int(x) for x in re.findall(r"ofSolve fix iterations:(\d+)", log)]
        errs = [float(x) for x in re.findall(r"Current error, s\d+ \d+ = ([0-9.eE+-]+)", log)]
        errs += [float(x) for x in re.findall(r"Final error=([0-9.eE+-]+)", log)]
        out["ms_%s_vel" % name] = v
        out["ms_%s_iters" % name] = np.array(iters, np.int32)
        out["ms_%s_errs" % name] = np.array(errs, np.float32)
        out["ms_%s_adv" % name] = ref.advect_semi_lagrange_cfl3d(999., v, a)
        print(name, dims, "iters", iters, "errs", errs, "max|v| %.3f" % np.abs(v).max())
    np.savez_compressed(os.path.join(HERE, "dim3_ops.npz"), **out)
    print("wrote dim3_ops.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
