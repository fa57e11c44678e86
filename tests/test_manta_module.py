"""The `manta` module (C++ host layer over the C ABI) driven like the reference's own grid tests
(tools/tests/test_0032_grid4dop.py style): small grids, helper plugins compared with numpy restatements of
grid4d.cpp:419-452, test.cpp:199-219 and optflow4d.cpp:1762-1777.  Runs in a subprocess because importing
`manta` creates the process-wide device context."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

SNIPPET = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
from manta import *
rng = np.random.default_rng(7)
nx, ny, nz, nt = 12, 10, 9, 8
s4 = Solver(name="t4", gridSize=vec3(nx, ny, nz), dim=3, fourthDim=nt)
sh = (nt, nz, ny, nx)

def put(g, a):
    g.fromBytes(np.ascontiguousarray(a).tobytes())

def get(g, shape, dt=np.float32):
    return np.frombuffer(g.toNumpyBytes(), dtype=dt).reshape(shape)

# grid4dMaxDiff / Vec4 / Vec3 / Int  (ref grid4d.cpp:419-464)
a, b = s4.create(Grid4Real), s4.create(Grid4Real)
A, B = rng.standard_normal(sh).astype(np.float32), rng.standard_normal(sh).astype(np.float32)
put(a, A); put(b, B)
assert abs(grid4dMaxDiff(a, b) - np.abs(A - B).max()) < 1e-6
v1, v2 = s4.create(Grid4Vec4), s4.create(Grid4Vec4)
V1, V2 = rng.standard_normal(sh + (4,)).astype(np.float32), rng.standard_normal(sh + (4,)).astype(np.float32)
put(v1, V1); put(v2, V2)
ref4 = np.abs(V1.astype(np.float64) - V2).sum(-1).max()
assert abs(grid4dMaxDiffVec4(v1, v2) - ref4) < 1e-5 * ref4
w1, w2 = s4.create(Grid4Vec3), s4.create(Grid4Vec3)
W1, W2 = rng.standard_normal(sh + (3,)).astype(np.float32), rng.standard_normal(sh + (3,)).astype(np.float32)
put(w1, W1); put(w2, W2)
ref3 = np.abs(W1.astype(np.float64) - W2).sum(-1).max()
assert abs(grid4dMaxDiffVec3(w1, w2) - ref3) < 1e-5 * ref3
i1, i2 = s4.create(Grid4Int), s4.create(Grid4Int)
I1, I2 = rng.integers(-50, 50, sh).astype(np.int32), rng.integers(-50, 50, sh).astype(np.int32)
put(i1, I1); put(i2, I2)
assert grid4dMaxDiffInt(i1, i2) == float(np.abs(I1.astype(np.int64) - I2).max())
assert grid4dMaxDiffInt(i1, i1) == 0.0

# debugGridAvg4d / debugVelAvg4d  (ref test.cpp:199-219)
for brd in (0, 2):
    sl = (slice(brd, nt - brd), slice(brd, nz - brd), slice(brd, ny - brd), slice(brd, nx - brd))
    refa = A[sl].astype(np.float64).mean() * 1e6
    assert abs(debugGridAvg4d(a, brd) - refa) <= 1e-4 * max(1.0, abs(refa))
    refv = np.sqrt((V1[sl].astype(np.float32) ** 2).sum(-1, dtype=np.float32)).astype(np.float64).mean()
    assert abs(debugVelAvg4d(v1, brd) - refv) <= 1e-5 * refv

# calcObfDiff  (ref optflow4d.cpp:1762-1777, 3D)
s3 = Solver(name="t3", gridSize=vec3(nx, ny, nz), dim=3)
sh3 = (nz, ny, nx)
p1, p2, pd, vd = s3.create(RealGrid), s3.create(RealGrid), s3.create(RealGrid), s3.create(RealGrid)
u1, u2, t1, t2 = s3.create(VecGrid), s3.create(VecGrid), s3.create(RealGrid), s3.create(RealGrid)
P1, P2 = rng.standard_normal(sh3).astype(np.float32), rng.standard_normal(sh3).astype(np.float32)
U1, U2 = rng.standard_normal(sh3 + (3,)).astype(np.float32), rng.standard_normal(sh3 + (3,)).astype(np.float32)
T1, T2 = rng.standard_normal(sh3).astype(np.float32), rng.standard_normal(sh3).astype(np.float32)
for g, x in ((p1, P1), (p2, P2), (u1, U1), (u2, U2), (t1, T1), (t2, T2)):
    put(g, x)
pd.setConst(-7.); vd.setConst(-7.)
bnd = 1
calcObfDiff(p1, p2, pd, u1, u2, t1, t2, vd, bnd)
PD, VD = get(pd, sh3), get(vd, sh3)
inner = (slice(bnd, nz - bnd), slice(bnd, ny - bnd), slice(bnd, nx - bnd))
assert np.array_equal(PD[inner], np.abs(P1 - P2)[inner])
d = np.concatenate([U1 - U2, (T1 - T2)[..., None]], -1)
refn = np.sqrt((d * d).sum(-1, dtype=np.float32))
assert np.allclose(VD[inner], refn[inner], rtol=2e-7, atol=0)
mask = np.ones(sh3, bool); mask[inner] = False
assert (PD[mask] == -7.).all() and (VD[mask] == -7.).all()   # FOR_IJK_BND leaves the border untouched
print("MANTA_HELPERS_OK")
"""


def test_helper_plugins_match_reference_semantics():
    host = os.path.join(ROOT, "ofblend_b200", "host")
    p = subprocess.run([sys.executable, "-c", SNIPPET, host], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0 and "MANTA_HELPERS_OK" in p.stdout, p.stdout[-2000:] + "\n" + p.stderr[-3000:]


DEFOVOL_SNIPPET = r"""
import os, sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
tmp = sys.argv[3]
from manta import *
from oracle import port   # checker only
rng = np.random.default_rng(5)
dd = (8, 7, 8, 20)
big = (30, 28, 30, 30)

def put(g, a):
    g.fromBytes(np.ascontiguousarray(a).tobytes())

def get(g, shape, dt=np.float32):
    return np.frombuffer(g.toNumpyBytes(), dtype=dt).reshape(shape)

sd = Solver(name="defo", gridSize=vec3(dd[0], dd[1], dd[2]), dim=3, fourthDim=dd[3])
vols, fns = [], []
for q in range(3):
    V = (rng.standard_normal((dd[3], dd[2], dd[1], dd[0], 4)) * 0.5).astype(np.float32)
    g = sd.create(Grid4Vec4); put(g, V)
    fn = os.path.join(tmp, "defo%d.uni" % q); g.save(fn)
    vols.append(V); fns.append(fn)
flushUniWrites()
sb = Solver(name="big", gridSize=vec3(big[0], big[1], big[2]), dim=3, fourthDim=big[3])
s3 = Solver(name="out", gridSize=vec3(big[0], big[1], big[2]), dim=3)
PHI = rng.standard_normal((big[3], big[2], big[1], big[0])).astype(np.float32)
phi = sb.create(Grid4Real); put(phi, PHI)
dst = s3.create(RealGrid)
fac = vec4(*[big[i] / dd[i] for i in range(4)])
facn = tuple(big[i] / dd[i] for i in range(4))
osz = vec4(*[float(x) for x in big])
times = [0.9, 2.4, 3.6, 14.5, 16., 29.]
sh3 = (big[2], big[1], big[0])
for n, aligned in ((2, False), (2, True), (3, False)):
    loadAdvectTimeSlice_OptInit(7, fns[0], True, aligned, 0.1)
    for q in range(1, n):
        loadAdvectTimeSlice_OptAdd(7, fns[q])
    want = port.load_advect_defovols(vols[:n], big[:3], PHI, times, 0.6, 0.3, 0.2, 1., 0., 1., facn, doAligned=aligned,
                                     partialLoadFac=0.1, overrideSize=big, overrideTimeOff=0.5, bordSkip=3, defoAniFac=0.75)
    for f, tm in enumerate(times):
        dst.setConst(0.)
        loadAdvectTimeSlice_OptRun(7, fns[0], dst, phi, tm, 0.6, 1., vec4(0.), vec4(1.), fac, overrideSize=osz, overrideTimeOff=0.5,
                                   thirdAlpha=0.3, bordSkip=3, fourthAlpha=0.2, defoAniFac=0.75)
        got = get(dst, sh3)
        assert np.array_equal(got, want[f]), (n, aligned, f, np.abs(got - want[f]).max())
    loadAdvectTimeSlice_Finish(7)
# the unoptimised twin with its debug outputs
dv, dt = s3.create(VecGrid), s3.create(RealGrid)
dst.setConst(0.)
loadAdvectTimeSlice(0, fns[0], dst, phi, 6.5, 0.7, 1., vec4(0.), vec4(1.), fac, overrideSize=osz, overrideTimeOff=-0.5, debugVel=dv, debugVelT=dt,
                    defoAniFac=0.8)
w = port.load_advect_time_slice_unopt(vols[0], big[:3], PHI, 6.5, 0.7, 1., 0., 1., facn, big, -0.5, 0.8, False)
assert np.array_equal(get(dst, sh3), w[0]) and np.abs(w[0]).max() > 0
assert np.array_equal(get(dv, sh3 + (3,)), w[1]) and np.array_equal(get(dt, sh3), w[2])
print("MANTA_DEFOVOL_OK")
"""


def test_defo_volumes_and_unoptimised_lookup_through_the_module(tmp_path):
    """`thirdload` through the plugin API exactly as flof.py would drive it (ref flof.py:757-816): _OptInit(useDefoVols=True),
    _OptAdd, _OptRun with thirdAlpha / fourthAlpha, _Finish; files written by the module itself; frames compared bit for bit
    with the oracle.  Also loadAdvectTimeSlice with debugVel / debugVelT."""
    host = os.path.join(ROOT, "ofblend_b200", "host")
    p = subprocess.run([sys.executable, "-c", DEFOVOL_SNIPPET, host, ROOT, str(tmp_path)], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0 and "MANTA_DEFOVOL_OK" in p.stdout, p.stdout[-2000:] + "\n" + p.stderr[-3000:]


DIM3_SNIPPET = r"""
import os, sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2]); sys.path.insert(0, os.path.join(sys.argv[2], "tests"))
from manta import *
from conftest import DIM3_CASES, sdf_pair3
gold = np.load(os.path.join(sys.argv[2], "tests", "golden", "dim3_ops.npz"))

def put(g, a):
    g.fromBytes(np.ascontiguousarray(a).tobytes())

def get(g, shape, dt=np.float32):
    return np.frombuffer(g.toNumpyBytes(), dtype=dt).reshape(shape)

# the call sequence of scenes/opticalFlowSimple3d.py:12-64 (inputs uploaded instead of Box.computeLevelset)
dims, params = DIM3_CASES["scene"]
A, B = sdf_pair3(dims)
gs = vec3(*dims)
s = Solver(name='main', gridSize=gs, dim=3)
flags = s.create(FlagGrid)
phi, i0, i1 = s.create(LevelsetGrid), s.create(LevelsetGrid), s.create(LevelsetGrid)
vel = s.create(MACGrid)
put(i0, A); put(i1, B)
opticalFlowMultiscale3d(i0=i0, i1=i1, vel=vel, wSmooth=0.5, wEnergy=0.0001, multiStep=4)
phi.copyFrom(i0)
advectSemiLagrangeCfl(flags=flags, vel=vel, grid=phi, order=1, velFactor=(float)(1.0), cfl=999)
sh = A.shape
assert np.array_equal(get(vel, sh + (3,)), gold["ms_scene_vel"])
assert np.array_equal(get(phi, sh), gold["ms_scene_adv"])
# (bnd = 2: the advection leaves a zero shell, ref :820-834)
e0, e1 = calcLsDiff3d(i0=i0, i1=i1, correction=20., bnd=2), calcLsDiff3d(i0=phi, i1=i1, correction=20., bnd=2)
assert e1 < 0.5 * e0, (e0, e1)      # the deformation brings i0 onto i1
# advectCent3d = one semi-Lagrangian step (ref :836-846)
p2 = s.create(LevelsetGrid); p2.copyFrom(i0)
advectCent3d(vel, p2)
assert np.array_equal(get(p2, sh), gold["ms_scene_adv"])
# corrVelsOf3d on the second fixture
D = (20, 18, 16); SH = (16, 18, 20)
s2 = Solver(name='c', gridSize=vec3(*D), dim=3)
a0, a1 = sdf_pair3(D)
dst, v, q0, q1 = s2.create(VecGrid), s2.create(VecGrid), s2.create(RealGrid), s2.create(RealGrid)
put(q0, a0); put(q1, a1)
put(v, (np.random.default_rng(12).standard_normal(SH + (3,)) * 0.5).astype(np.float32))
corrVelsOf3d(dst, v, q0, q0, q1, 4., 1e10, 2., 0.1, 40)
assert np.array_equal(get(dst, SH + (3,)), gold["corr_dst"]) and np.array_equal(get(v, SH + (3,)), gold["corr_vel"])
try:
    opticalFlowMultiscale3d(i0=i0, i1=q1, vel=vel)
    raise SystemExit("size mismatch not detected")
except RuntimeError:
    pass
print("MANTA_DIM3_OK")
"""


def test_3d_plugins_through_the_module():
    """opticalFlowMultiscale3d / advectSemiLagrangeCfl / advectCent3d / calcLsDiff3d / corrVelsOf3d called like
    scenes/opticalFlowSimple3d.py does, results bit-identical to the reference's (tests/golden/dim3_ops.npz)."""
    host = os.path.join(ROOT, "ofblend_b200", "host")
    p = subprocess.run([sys.executable, "-c", DIM3_SNIPPET, host, ROOT], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0 and "MANTA_DIM3_OK" in p.stdout, p.stdout[-2000:] + "\n" + p.stderr[-3000:]
