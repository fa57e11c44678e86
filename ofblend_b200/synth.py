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


def two_drop_hires_slice(n, dataset, frame, nframes):
    """One hi-res 3D SDF frame (n^3, cell units, UNscaled and without border) of data set 0/1 -- the
    analogue of the reference's outxl_r080_x00N_%04d.uni files (scenes/dataGen2Drop.py:270-310)."""
    x = ((np.arange(n) + 0.5) / n)[None, None, :]
    y = ((np.arange(n) + 0.5) / n)[None, :, None]
    z = ((np.arange(n) + 0.5) / n)[:, None, None]
    tau = frame / float(max(nframes - 1, 1))
    g = (0.7 - 0.15) / (0.5 * 0.6 ** 2)
    y1 = max(0.7 - 0.5 * g * tau ** 2, 0.15)
    y2 = max(0.4 - 0.5 * g * tau ** 2, 0.15)
    basin = y - 0.15 + 0 * x + 0 * z
    if dataset == 0:
        a = _sphere(x, y, z, (0.25, y1, 0.33), 0.12)
        b = _box(x, y, z, (0.75, y2, 0.66), 0.08)
    else:
        a = _box(x, y, z, (0.45, y1, 0.33), 0.08)
        b = _sphere(x, y, z, (0.55, y2, 0.66), 0.12)
    return np.clip(np.minimum(basin, np.minimum(a, b)) * n, -10.0, 10.0).astype(np.float32)


def two_drop_4d_raw(n, nt, dataset):
    """Raw (unprocessed) 4D SDF n^3 x nt of data set 0/1 -- the analogue of out_r040_x00N.uni."""
    return np.stack([two_drop_hires_slice(n, dataset, t, nt) for t in range(nt)], axis=0)


def write_scene_inputs(outdir, res_load=40, res_xl=40, n_slices=181):
    """Writes the input files scenes/flof.py expects (setup 1, scenes/flof.py:123-150) from the analytic
    two-drop data: out_r040_x00{0,1}.uni (4D, res_load^3 x 1.5*res_load) and
    outxl_r080_x00{0,1}_%04d.uni (3D slices; flof.py reads their size from the header, so smaller
    slices than 80^3 are legal and keep the test light)."""
    import os
    from . import uni
    nt = int(res_load * 1.5)
    for ds in (0, 1):
        uni.write_uni(os.path.join(outdir, "out_r%03d_x%03d.uni" % (res_load, ds)), two_drop_4d_raw(res_load, nt, ds))
        for f in range(n_slices):
            uni.write_uni(os.path.join(outdir, "outxl_r%03d_x%03d_%04d.uni" % (2 * res_load, ds, f)),
                          two_drop_hires_slice(res_xl, ds, f, n_slices))


def analytic_deformation(dims, seed_phase=0.0, amp=(2.5, 2.0, 1.5, 3.0)):
    """Smooth analytic Vec4 deformation on an (nx, ny, nz, nt) grid, zero towards the border (like a
    mode-1 result after the border reset).  numpy only uses +,-,*,/ and np.sin here; the field is
    written to a .uni file once and both implementations read that file, so libm differences between
    machines cannot enter a comparison."""
    nx, ny, nz, nt = [int(d) for d in dims]
    x = ((np.arange(nx) + 0.5) / nx)[None, None, None, :]
    y = ((np.arange(ny) + 0.5) / ny)[None, None, :, None]
    z = ((np.arange(nz) + 0.5) / nz)[None, :, None, None]
    t = ((np.arange(nt) + 0.5) / nt)[:, None, None, None]
    win = (np.sin(np.pi * x) * np.sin(np.pi * y) * np.sin(np.pi * z) * np.sin(np.pi * t)) ** 2
    v = np.empty((nt, nz, ny, nx, 4), np.float32)
    v[..., 0] = amp[0] * win * np.sin(2 * np.pi * (y + 0.5 * t) + seed_phase)
    v[..., 1] = amp[1] * win * np.cos(2 * np.pi * (x - 0.3 * z) + seed_phase)
    v[..., 2] = amp[2] * win * np.sin(2 * np.pi * (x + y) - seed_phase)
    v[..., 3] = amp[3] * win * np.cos(2 * np.pi * (z + 0.25 * t) + seed_phase)
    return v
