"""End-to-end goldens of the drop-in boundary, produced by the REFERENCE's own `manta` executable
(oracle/_ref/manta, built by `make -C oracle refpy`) running the UNMODIFIED scenes/flof.py.

Dev container only.  Procedure:
  A  /tmp/scenesyn   analytic two-drop inputs (ofblend_b200.synth.write_scene_inputs), reference runs
                     `flof.py dataid0 0 dataid1 1 mode 1` and the reverse direction
  B  /tmp/scenesyn2  same inputs + ANALYTIC deformation files (synth.analytic_deformation), reference runs
                     mode 2 (one-way and two-way) and mode 3 (two-way alpha 50; partial-load streaming)
so that modes 2/3 are compared on bit-identical deformation input.
Stores sub-sampled reference outputs in tests/golden/scene_flof.npz.
    python tests/golden/make_scene_golden.py /tmp/scenesyn /tmp/scenesyn2
README configuration (BASELINE.json configs[0..2]; inputs from the reference's dataGen2Drop.py px {0,1}):
    python tests/golden/make_scene_golden.py /tmp/flofdata /tmp/flofdata scene_readme.npz
"""
import glob
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from ofblend_b200 import uni  # noqa: E402


def trace(logfile):
    log = open(logfile).read()
    iters = [int(x) for x in re.findall(r"ofSolve fix iterations:(\d+)", log)]
    errs = [float(x) for x in re.findall(r"Current error, s\d+ \d+ = ([0-9.eE+-]+)", log)]
    errs += [float(x) for x in re.findall(r"Final error=([0-9.eE+-]+)", log)]
    inp = [float(x) for x in re.findall(r"Error between inputs ([0-9.eE+-]+)", log)]
    return np.array(iters, np.int32), np.array(errs, np.float32), np.array(inp, np.float32)


def frames(d, prefix, pick):
    files = sorted(glob.glob(os.path.join(d, prefix + "_[0-9][0-9][0-9][0-9].uni")))
    nums = [int(f[-8:-4]) for f in files]
    out = {}
    for p in pick:
        n = nums[int(p * (len(nums) - 1))]
        out[n] = uni.read_uni(os.path.join(d, "%s_%04d.uni" % (prefix, n)))
    return nums, out


def main(da, db, outname="scene_flof.npz"):
    g = {}
    for tag, fn, log in (("01", "defo01_000_001_032_vel.uni", "ref_mode1_01.log"), ("10", "defo01_001_000_032_vel.uni", "ref_mode1_10.log")):
        v = uni.read_uni(os.path.join(da, fn))
        it, er, inp = trace(os.path.join(da, log))
        g["m1_%s_vel_sub" % tag] = v[::2, ::2, ::2, ::2].copy()
        g["m1_%s_vel_l2" % tag] = np.float64(np.linalg.norm(v.astype(np.float64).ravel()))
        g["m1_%s_iters" % tag] = it
        g["m1_%s_errs" % tag] = er
        g["m1_%s_input_err" % tag] = inp
    for tag, prefix in (("m2", "out_f0t1_a100"), ("m2tw", "out_f0t1_a030"), ("m3", "out_f0t1_a050")):
        if not glob.glob(os.path.join(db, prefix + "_[0-9][0-9][0-9][0-9].uni")):
            continue
        nums, fr = frames(db, prefix, (0.0, 0.35, 0.7, 1.0))
        g["%s_frame_numbers" % tag] = np.array(nums, np.int32)
        for n, a in fr.items():
            g["%s_frame_%04d" % (tag, n)] = a if a.size <= 70000 else a[::2, ::2, ::2].copy()
    np.savez_compressed(os.path.join(HERE, outname), **g)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in g.items()})


if __name__ == "__main__":
    main(*sys.argv[1:4])
