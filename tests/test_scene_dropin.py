"""Drop-in test of the boundary: the reference's UNMODIFIED scenes/flof.py (+ ofHelpers.py) is executed
against the B200 `manta` module (ofblend_b200/host, C++ over the C ABI) for modes 1, 2 and 3, and its
output files are compared with goldens that the reference's own `manta` executable wrote for the same
input files (tests/golden/make_scene_golden.py).

The scene scripts are not part of this repository: `make -C oracle ref` copies them byte for byte into
the git-ignored oracle/_ref/scenes/, which travels to the GPU box like the built .so files."""
import glob
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENE = os.path.join(ROOT, "oracle", "_ref", "scenes", "flof.py")
GOLD = os.path.join(ROOT, "tests", "golden", "scene_flof.npz")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.isfile(SCENE), reason="oracle/_ref/scenes/flof.py missing (run `make -C oracle ref`)"),
              pytest.mark.skipif(not os.path.isfile(GOLD), reason="tests/golden/scene_flof.npz missing")]


def run_flof(cwd, *args):
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    p = subprocess.run([sys.executable, "-m", "ofblend_b200.run_scene", SCENE] + [str(a) for a in args], cwd=cwd, env=env,
                       capture_output=True, text=True, timeout=1200)
    assert p.returncode == 0, p.stdout[-3000:] + "\n" + p.stderr[-3000:]
    return p.stdout


@pytest.fixture(scope="module")
def inputs(tmp_path_factory):
    from ofblend_b200 import synth
    d = str(tmp_path_factory.mktemp("scene"))
    synth.write_scene_inputs(d)
    return d


def test_mode1_scene(inputs):
    from ofblend_b200 import uni
    g = np.load(GOLD)
    for tag, a0, a1, fn in (("01", 0, 1, "defo01_000_001_032_vel.uni"), ("10", 1, 0, "defo01_001_000_032_vel.uni")):
        out = run_flof(inputs, "dataid0", a0, "dataid1", a1, "mode", 1)
        iters = [int(x) for x in re.findall(r"ofSolve fix iterations:(\d+)", out)]
        errs = [float(x) for x in re.findall(r"Current error, step \d+ = ([0-9.eE+-]+)", out)]
        inp = [float(x) for x in re.findall(r"Error between inputs ([0-9.eE+-]+)", out)]
        assert iters == [int(x) for x in g["m1_%s_iters" % tag]], (iters, g["m1_%s_iters" % tag])
        assert np.allclose(inp, g["m1_%s_input_err" % tag], rtol=1e-5)          # load + pre-processing path
        assert np.allclose(errs, g["m1_%s_errs" % tag][:-1], rtol=1e-4), (errs, g["m1_%s_errs" % tag])
        vel = uni.read_uni(os.path.join(inputs, fn))
        assert vel.shape == (48, 32, 32, 32, 4)
        # the deformation file the product writes (after the final SDF projection) equals the reference's bit for bit:
        # the CG dot products are summed in the reference's sequential order
        assert np.array_equal(vel[::2, ::2, ::2, ::2], g["m1_%s_vel_sub" % tag]), rel_l2(vel[::2, ::2, ::2, ::2], g["m1_%s_vel_sub" % tag])
        assert abs(np.linalg.norm(vel.astype(np.float64).ravel()) - float(g["m1_%s_vel_l2" % tag])) <= 1e-12 * float(g["m1_%s_vel_l2" % tag])


def _check_frames(d, g, tag, prefix, tol_cells):
    from ofblend_b200 import uni
    files = sorted(glob.glob(os.path.join(d, prefix + "_[0-9][0-9][0-9][0-9].uni")))
    nums = [int(f[-8:-4]) for f in files]
    assert nums == [int(x) for x in g["%s_frame_numbers" % tag]], (nums[:4], g["%s_frame_numbers" % tag][:4])
    checked = 0
    worst = 0.0
    for key in g.files:
        m = re.match(r"%s_frame_(\d{4})$" % tag, key)
        if not m:
            continue
        a = uni.read_uni(os.path.join(d, "%s_%s.uni" % (prefix, m.group(1))))
        if a.size > 70000:
            a = a[::2, ::2, ::2]
        worst = max(worst, float(np.abs(a - g[key]).max()))
        assert np.abs(a - g[key]).max() <= tol_cells, (key, np.abs(a - g[key]).max())
        checked += 1
    assert checked >= 3
    print("%s: %d frames checked, max-abs difference to the reference chain %.3g cells" % (tag, checked, worst))
    return worst


def test_mode2_mode3_scene(inputs):
    """Modes 2 and 3 on bit-identical (analytic) deformation files: applied SDF within 1e-3 cells."""
    from ofblend_b200 import synth, uni
    g = np.load(GOLD)
    uni.write_uni(os.path.join(inputs, "defo01_000_001_032_vel.uni"), synth.analytic_deformation((32, 32, 32, 48), 0.0))
    uni.write_uni(os.path.join(inputs, "defo01_001_000_032_vel.uni"),
                  synth.analytic_deformation((32, 32, 32, 48), 1.3, amp=(-2.0, 1.5, -2.5, -3.0)))
    for f in glob.glob(os.path.join(inputs, "*_prep*.uni")):
        os.remove(f)
    run_flof(inputs, "dataid0", 0, "dataid1", 1, "mode", 2, "writeuni", 1)
    _check_frames(inputs, g, "m2", "out_f0t1_a100", 1e-3)
    run_flof(inputs, "dataid0", 0, "dataid1", 1, "mode", 2, "twoway", 1, "alpha", 30, "writeuni", 1)
    _check_frames(inputs, g, "m2tw", "out_f0t1_a030", 1e-3)
    run_flof(inputs, "mode", 3, "twoway", 1, "alpha", 50, "writeuni", 1)
    _check_frames(inputs, g, "m3", "out_f0t1_a050", 1e-3)


README_SDF = os.path.join(ROOT, "tests", "golden", "readme")               # committed: the two 4D SDFs (40^3 x 60)
README_HIRES = os.path.join(ROOT, "oracle", "_ref", "data", "readme")       # built: 2 x 181 hi-res slices (80^3), mode 3 only
README_GOLD = os.path.join(ROOT, "tests", "golden", "scene_readme.npz")


@pytest.fixture(scope="module")
def readme_mode1(tmp_path_factory):
    """BASELINE.json configs[0]: `flof.py dataid0 0 dataid1 1 mode 1` (+ the reverse direction) on the reference's own
    example data (dataGen2Drop.py px 0/1, res 40; tests/golden/readme/README.md says how it was generated)."""
    from ofblend_b200 import uni
    g = np.load(README_GOLD)
    d = str(tmp_path_factory.mktemp("readme"))
    for f in os.listdir(README_SDF):
        if f.endswith(".uni"):
            os.symlink(os.path.join(README_SDF, f), os.path.join(d, f))
    for tag, a0, a1, fn in (("01", 0, 1, "defo01_000_001_032_vel.uni"), ("10", 1, 0, "defo01_001_000_032_vel.uni")):
        out = run_flof(d, "dataid0", a0, "dataid1", a1, "mode", 1)
        iters = [int(x) for x in re.findall(r"ofSolve fix iterations:(\d+)", out)]
        errs = [float(x) for x in re.findall(r"Current error, step \d+ = ([0-9.eE+-]+)", out)]
        inp = [float(x) for x in re.findall(r"Error between inputs ([0-9.eE+-]+)", out)]
        assert iters == [int(x) for x in g["m1_%s_iters" % tag]], (iters, g["m1_%s_iters" % tag])   # 27,25,24,27,27,27
        assert np.allclose(inp, g["m1_%s_input_err" % tag], rtol=1e-5)                              # 350.4529
        assert np.allclose(errs, g["m1_%s_errs" % tag][:-1], rtol=1e-5), (errs, g["m1_%s_errs" % tag])
        vel = uni.read_uni(os.path.join(d, fn))
        # the product's deformation file (32^3 x 48, after the final projection) == the reference's, bit for bit
        assert np.array_equal(vel[::2, ::2, ::2, ::2], g["m1_%s_vel_sub" % tag]), rel_l2(vel[::2, ::2, ::2, ::2], g["m1_%s_vel_sub" % tag])
        assert abs(np.linalg.norm(vel.astype(np.float64).ravel()) - float(g["m1_%s_vel_l2" % tag])) <= 1e-12 * float(g["m1_%s_vel_l2" % tag])
    return d


def test_readme_mode1_then_mode2(readme_mode1):
    """configs[0] and [1]: mode 2 applied with the deformation the PRODUCT computed (chain GPU mode 1 -> GPU mode 2)
    against the frames of the reference's own chain: applied SDF within 1e-3 cells (measured: identical)."""
    g = np.load(README_GOLD)
    run_flof(readme_mode1, "dataid0", 0, "dataid1", 1, "mode", 2, "writeuni", 1)
    _check_frames(readme_mode1, g, "m2", "out_f0t1_a100", 1e-3)


@pytest.mark.skipif(not glob.glob(os.path.join(README_HIRES, "outxl_r080_x001_*.uni")),
                    reason="hi-res slices of the README data set missing (oracle/_ref/data/readme, `make -C oracle refdata`)")
def test_readme_mode3(readme_mode1):
    """configs[2]: `flof.py mode 3 twoway 1 alpha 50` -- two-way blended application of the product's own two
    deformations to the full-resolution slice sequence, against the reference chain's frames."""
    g = np.load(README_GOLD)
    for f in os.listdir(README_HIRES):
        if f.startswith("outxl_") and not os.path.exists(os.path.join(readme_mode1, f)):
            os.symlink(os.path.join(README_HIRES, f), os.path.join(readme_mode1, f))
    run_flof(readme_mode1, "mode", 3, "twoway", 1, "alpha", 50, "writeuni", 1)
    _check_frames(readme_mode1, g, "m3", "out_f0t1_a050", 1e-3)
