import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/libofref.so (the compiled reference)")


def pytest_collection_modifyitems(config, items):
    from oracle import ref
    if not ref.available():
        skip = pytest.mark.skip(reason="oracle/_ref/libofref.so not built (needs /root/reference)")
        for it in items:
            if "ref" in it.keywords:
                it.add_marker(skip)


def sdf_pair(dims, seed=0):
    """Small smooth SDF-like test pair (two shifted blobs), float32 [t,z,y,x], values ~[-0.2,0.2]."""
    nx, ny, nz, nt = dims
    t, z, y, x = np.meshgrid(np.arange(nt), np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    rng = np.random.default_rng(seed)

    def blob(cx, cy, cz, ct, r):
        return np.sqrt((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2 + 0.5 * (t - ct) ** 2) - r

    c = np.array([nx, ny, nz, nt]) / 2.0
    a = blob(c[0] - 1.2, c[1], c[2] + 0.5, c[3], 0.28 * nx)
    b = blob(c[0] + 1.0, c[1] + 0.8, c[2], c[3] - 0.6, 0.30 * nx)
    a = a + 0.05 * rng.standard_normal(a.shape)
    b = b + 0.05 * rng.standard_normal(b.shape)
    return (np.clip(a, -10, 10) * -0.005).astype(np.float32), (np.clip(b, -10, 10) * -0.005).astype(np.float32)


def sdf_pair3(dims, seed=0):
    """3D test pair for the 3D instantiations: two shifted blobs, float32 [z,y,x], SDF in cells * -0.005."""
    nx, ny, nz = dims
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    rng = np.random.default_rng(seed + 7)

    def blob(cx, cy, cz, r):
        return np.sqrt((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2) - r

    c = np.array([nx, ny, nz]) / 2.0
    a = blob(c[0] - 1.2, c[1], c[2] + 0.5, 0.28 * nx) + 0.05 * rng.standard_normal(z.shape)
    b = blob(c[0] + 1.0, c[1] + 0.8, c[2], 0.30 * nx) + 0.05 * rng.standard_normal(z.shape)
    return (np.clip(a, -10, 10) * -0.005).astype(np.float32), (np.clip(b, -10, 10) * -0.005).astype(np.float32)


# opticalFlowMultiscale3d cases of the golden fixture (tests/golden/make_golden3d.py): dims, keyword arguments
DIM3_CASES = {
    # the parameters of scenes/opticalFlowSimple3d.py:60
    "scene": ((24, 24, 24), dict(wSmooth=0.5, wEnergy=0.0001, multiStep=4)),
    # the flof.py parameter set (blur, border reset, two levels, final projection) on a 3D pair
    "flof": ((32, 32, 32), dict(wSmooth=1e-3, wEnergy=1e-4, postVelBlur=4., cgAccuracy=1e-2, resetBndWidth=0.1, multiStep=3,
                                minGridSize=20, doFinalProject=True)),
    # a 2D grid (nz == 1): the DIM = 2 instantiation of scenes/ofblend2dTest.py, three levels, final projection
    "plane": ((40, 36, 1), dict(wSmooth=1e-2, wEnergy=1e-4, postVelBlur=4., cgAccuracy=1e-3, resetBndWidth=0.1, multiStep=3,
                                minGridSize=12, doFinalProject=True)),
    # projection instead of the solve on the fine level (projSizeThresh), non-cubic grid
    "proj": ((28, 24, 20), dict(wSmooth=1e-2, wEnergy=1e-4, postVelBlur=2., cgAccuracy=1e-3, resetBndWidth=0.1, multiStep=2,
                                minGridSize=12, projSizeThresh=20)),
}


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))
