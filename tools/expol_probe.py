"""Measures what the extrapolation sweeps of corrVelsOf4d actually work on (marker density, how the
non-zero region grows per sweep) on the synthetic pair:  python tools/expol_probe.py 64"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ofblend_b200 import capi, synth  # noqa: E402


def main():
    res = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dims = (res, res, res, res)
    ctx = capi.Context(0)
    api = capi.HostAPI(ctx)
    i0 = ctx.to_device(synth.post_process(synth.two_drop_phi(dims, 0), api))
    i1 = ctx.to_device(synth.post_process(synth.two_drop_phi(dims, 1), api))
    vel = ctx.grid(dims, 4)
    dst = ctx.grid(dims, 4)
    mk = ctx.grid(dims, 1)
    P = capi.make_params(wSmooth=1e-3, wEnergy=1e-4, postVelBlur=4., cgAccuracy=1e-2, resetBndWidth=0.1, multiStep=3,
                         minGridSize=20, doFinalProject=False)
    ctx.optical_flow_multiscale4d(vel, i0, i1, P)
    lo, hi, _ = ctx.grid_min_max(vel)
    print("max |vel| %.3f -> sweeps %d" % (hi, int(hi + 4)))
    ctx.project_cells(dst, vel, i0, i1, mk, 4., 40)
    m = mk.download().reshape(dims[::-1])
    v = dst.download().reshape(dims[::-1] + (4,))
    interior = np.zeros(m.shape, bool)
    interior[1:-1, 1:-1, 1:-1, 1:-1] = True
    un = (m == 0) & interior
    print("cells %d  unmarked interior %.4f  marked %.4f" % (m.size, un.mean(), (m != 0).mean()))
    # patches of 32 x * 4 y
    p = un.reshape(res, res, res // 4, 4, res // 32, 32).any(axis=(3, 5))
    print("active 32x4 patches %.4f" % p.mean())
    p2 = un.reshape(res, res, res // 8, 8, res // 32, 32).any(axis=(3, 5))
    print("active 32x8 patches %.4f" % p2.mean())
    un2 = un.copy()
    un2[..., 1] = False
    un2[..., res - 2] = False
    print("unmarked interior without the x shell %.4f" % un2.mean())
    for py, pz in ((4, 1), (8, 1), (4, 2), (4, 4), (2, 2), (8, 2)):
        q = un2.reshape(res, res // pz, pz, res // py, py, res // 32, 32)
        print("active 32x%dx%d warp patches (no x shell) %.4f   lanes used in active patches %.4f   thread patches 1x%dx%d active %.4f" % (
            py, pz, q.any(axis=(2, 4, 6)).mean(), q.any(axis=(2, 4)).sum() / max(1, q.any(axis=(2, 4, 6)).sum() * 32), py, pz,
            q.any(axis=(2, 4)).mean()))
    nz = (v != 0).any(axis=-1)
    for s in (0, 1, 2, 4, 8, 16, 32, 64):
        if s:
            ctx.cv_expol_blur4d(dst, mk, s - prev)
        prev = s
        v = dst.download().reshape(dims[::-1] + (4,))
        nz = (v != 0).any(axis=-1)
        print("after %2d sweeps: nonzero cells %.4f   unmarked&nonzero %.4f of unmarked" % (s, nz.mean(), (nz & un).sum() / un.sum()))
    ctx.close()


if __name__ == "__main__":
    main()
