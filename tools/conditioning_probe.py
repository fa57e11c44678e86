"""How well conditioned is the reference's mode-1 result?  (CPU only, uses the oracle; ~3 min)

Runs the C oracle (bit-identical to the reference in dot mode 0) on the synthetic 32^3x48 pair twice:
with the reference's sequential fp64 dot-product summation and with the same products summed in
blocks of 4096, with and without the final SDF projection, and prints how far the deformation moves.
Result recorded in DESIGN.md §2: 8e-8 rel-L2 before the projection, 3.0e-3 after it."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import port  # noqa: E402
from ofblend_b200 import synth  # noqa: E402


def rel_l2(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


port.set_threads(os.cpu_count() or 1)
dims = (32, 32, 32, 48)
i0 = synth.post_process(synth.two_drop_phi(dims, 0), port)
i1 = synth.post_process(synth.two_drop_phi(dims, 1), port)
v0 = np.zeros(i0.shape + (4,), np.float32)
out = {}
for mode in (0, 1):
    port.lib().orc_set_dot_mode(mode)
    for proj in (True, False):
        p = dict(synth.MODE1_PARAMS)
        p["doFinalProject"] = proj
        t = time.time()
        out[(mode, proj)] = port.optical_flow_multiscale4d(v0, i0, i1, **p)
        print("dot mode %d, projection %s: %.1f s" % (mode, proj, time.time() - t), flush=True)
port.lib().orc_set_dot_mode(0)
for proj in (False, True):
    a, b = out[(1, proj)], out[(0, proj)]
    print("projection %-5s: blocked vs sequential summation  rel-L2 %.3g  max-abs %.3g cells  differing values %d"
          % (proj, rel_l2(a, b), np.abs(a - b).max(), int((a != b).sum())))
