"""Synthetic two-drop 4D SDF pairs (BASELINE.json configs 4/5).

Analytic, no RNG.  Geometry follows the reference data generator scenes/dataGen2Drop.py:115,
141-155 (basin + two falling drops, one sphere and one box, swapped between the two data
sets) placed in the inner 80 % of the domain like the loader of scenes/flof.py:536-556 does
with loadOffset / loadScale (scenes/ofHelpers.py:59-69).  phi is in cell units, negative
inside, clamped to +-10 (dataGen2Drop.py:290-293).

numpy layout everywhere in this package: array[t, z, y, x] (x fastest), Vec4 grids
array[t, z, y, x, 4] -- byte-identical to the reference's Grid4d<T> storage
(source/grid4d.h:92-97).
"""
import numpy as np

BORDER = 0.1           # ofsd["autoBorder"], scenes/flof.py:313
LS_FACTOR = -0.1 / 20  # ofsd["lsFactor"],  scenes/flof.py:319
MAX_DIST = 40          # ofsd["maxDist"],   scenes/flof.py:320


def _sphere(x, y, z, c, r):
    return np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) - r


def _box(x, y, z, c, h):
    qx = np.abs(x - c[0]) - h
    qy = np.abs(y - c[1]) - h
    qz = np.abs(z - c[2]) - h
    outside = np.sqrt(np.maximum(qx, 0) ** 2 + np.maximum(qy, 0) ** 2 + np.maximum(qz, 0) ** 2)
    inside = np.minimum(np.maximum(qx, np.maximum(qy, qz)), 0)
    return outside + inside


def two_drop_phi(dims, dataset):
    """Raw SDF of data set 0 or 1 on an (nx, ny, nz, nt) grid -> float32 [t, z, y, x]."""
    nx, ny, nz, nt = [int(d) for d in dims]
    sc = 1.0 - 2.0 * BORDER
    x = (((np.arange(nx) + 0.5) / nx - BORDER) / sc)[None, None, :]
    y = (((np.arange(ny) + 0.5) / ny - BORDER) / sc)[None, :, None]
    z = (((np.arange(nz) + 0.5) / nz - BORDER) / sc)[:, None, None]
    # first 10 % of the data's time range repeats the start frame (repeatStartDist)
    tau = np.clip(((np.arange(nt) + 0.5) / nt - 2 * BORDER) / (1.0 - 3 * BORDER), 0.0, 1.0)
    g = (0.7 - 0.15) / (0.5 * 0.6 ** 2)
    out = np.empty((nt, nz, ny, nx), np.float32)
    cell = sc * nx  # unit-cube length -> cells
    for t in range(nt):
        y1 = max(0.7 - 0.5 * g * tau[t] ** 2, 0.15)
        y2 = max(0.4 - 0.5 * g * tau[t] ** 2, 0.15)
        basin = y - 0.15 + 0 * x + 0 * z
        if dataset == 0:
            a = _sphere(x, y, z, (0.25, y1, 0.33), 0.12)
            b = _box(x, y, z, (0.75, y2, 0.66), 0.08)
        else:
            a = _box(x, y, z, (0.45, y1, 0.33), 0.08)
            b = _sphere(x, y, z, (0.55, y2, 0.66), 0.12)
        phi = np.minimum(basin, np.minimum(a, b)) * cell
        out[t] = np.clip(phi, -10.0, 10.0).astype(np.float32)
    return out


def post_process(phi, ops, res=None):
    """scenes/flof.py:491-520 postProcMode 1 with the README parameters: border reset,
    outside + inside extrapolation, SDF rescale.  `ops` supplies set_bound4d(a, value, w),
    extrap4d_ls_simple(a, distance, inside) and mult_const(a, s) -- the GPU product in
    bench.py, the oracle in CPU tests."""
    res = res or phi.shape[3]
    a = ops.set_bound4d(phi, 0.1, int(res * BORDER))
    a = ops.extrap4d_ls_simple(a, MAX_DIST, False)
    a = ops.extrap4d_ls_simple(a, MAX_DIST, True)
    return ops.mult_const(a, LS_FACTOR)


# README solver parameters, scenes/flof.py:304-320, 913-915
MODE1_PARAMS = dict(wSmooth=0.001, wEnergy=0.0001, cgAccuracy=0.01, postVelBlur=4.0, cfl=999.0,
                    multiStep=3, minGridSize=20, doFinalProject=True, resetBndWidth=0.1)
