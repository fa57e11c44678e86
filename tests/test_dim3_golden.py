"""The committed 3D fixture (tests/golden/dim3_ops.npz) against the live compiled reference: the GPU tests of
the 3D instantiations compare with this file, so it must be exactly what the reference computes from the seeded inputs."""
import os

import numpy as np
import pytest

from conftest import DIM3_CASES, sdf_pair3
from oracle import ref

pytestmark = pytest.mark.ref
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dim3_ops.npz"))


def rnd(shape, seed, scale=1.0):
    return (np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32)


def test_operator_fixtures_are_the_references_outputs():
    D = (20, 18, 16)
    SH = (D[2], D[1], D[0])
    i0, i1 = sdf_pair3(D)
    vel = rnd(SH + (3,), 3, 2.0)
    assert np.float32(ref.calc_ls_diff3d(i0, i1, 0.005 * 200, 0)) == GOLD["lsdiff"][0]
    assert np.array_equal(ref.advect_semi_lagrange_cfl3d(1.0, vel, i0), GOLD["adv_real_cfl1"])
    assert np.array_equal(ref.advect_semi_lagrange_cfl3d(999., vel, rnd(SH + (3,), 4)), GOLD["adv_vec3"])
    d, v = ref.corr_vels_of3d(np.zeros(SH + (3,), np.float32), rnd(SH + (3,), 12, 0.5), i0, i1, 4., 2., 0.1, 40)
    assert np.array_equal(d, GOLD["corr_dst"]) and np.array_equal(v, GOLD["corr_vel"])


@pytest.mark.parametrize("name", sorted(DIM3_CASES))
def test_multiscale3d_fixture_is_the_references_output(name):
    dims, params = DIM3_CASES[name]
    a, b = sdf_pair3(dims)
    v = ref.optical_flow_multiscale3d(np.zeros(a.shape + (3,), np.float32), a, b, **params)
    assert np.array_equal(v, GOLD["ms_%s_vel" % name])
    assert np.abs(v).max() > 0.5                                   # a real deformation, not a trivial fixture
    # serial and OpenMP runs of the reference agree bit for bit
    n = ref.set_threads(1)
    try:
        assert np.array_equal(ref.optical_flow_multiscale3d(np.zeros(a.shape + (3,), np.float32), a, b, **params), v)
    finally:
        ref.set_threads(os.cpu_count() or 1)
