"""Fill statistics of the extrapolation work (cells with marker == 0) for candidate tilings of the sweep kernel:
   python tools/expol_fill_probe.py 128"""
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
    ctx.project_cells(dst, vel, i0, i1, mk, 4., 40)
    m = mk.download().reshape(dims[::-1])          # [t, z, y, x]
    un = np.zeros(m.shape, bool)
    un[1:-1, 1:-1, 1:-1, 1:-1] = m[1:-1, 1:-1, 1:-1, 1:-1] == 0
    del m
    need = int(un.sum())
    print("res %d cells %d needed %d (%.4f)" % (res, un.size, need, need / un.size))
    for P in (2, 4, 8):  # how evenly the sweep work spreads over the t-slabs of P ranks
        per = un.reshape(P, res // P, -1).sum(axis=(1, 2)).astype(np.float64)
        print("P=%d: needed cells per t-slab / mean: %s  (max %.3f)" % (P, " ".join("%.2f" % x for x in per / per.mean()), per.max() / per.mean()))
    shell = un.copy()
    shell[2:-2, 2:-2, 2:-2, 2:-2] = False
    print("needed cells on the index-1 shell: %.4f of needed" % (shell.sum() / need))
    del shell
    # outputs computed by a tiling = (tiles with >= 1 needed cell) * cells per tile, fill = needed / computed
    for (sx, py, pz) in ((1, 4, 4), (1, 4, 2), (1, 2, 2), (1, 1, 1), (32, 1, 1), (64, 1, 1), (res, 1, 1), (32, 4, 4), (64, 4, 4), (res, 4, 4), (res, 2, 2), (res, 4, 2)):
        if res % sx or res % py or res % pz:
            continue
        q = un.reshape(res, res // pz, pz, res // py, py, res // sx, sx)
        tiles = q.any(axis=(2, 4, 6))
        ntile = int(tiles.sum())
        # row-exact: within an active tile only the (zo, oy) rows that hold a needed cell are computed
        rows = q.any(axis=6)                       # [t, zb, zo, yb, oy, xs]
        nrow = int(rows.sum())
        print("tile %3dx x %dy x %dz: active tiles %9d (%.3f of all)  fill %.3f | row-exact inside tile: computed %.3f of grid, fill %.3f" % (
            sx, py, pz, ntile, tiles.mean(), need / (ntile * sx * py * pz), nrow * sx / un.size, need / (nrow * sx)))
    ctx.close()


if __name__ == "__main__":
    main()
