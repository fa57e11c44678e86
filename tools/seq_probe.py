"""Statistics of the sequential-order dot products on one finest-level CG solve:  python tools/seq_probe.py RES"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ofblend_b200 import capi, synth  # noqa: E402

res = int(sys.argv[1])
dims = (res,) * 4
ctx = capi.Context(0)
api = capi.HostAPI(ctx)
i0 = ctx.to_device(synth.post_process(synth.two_drop_phi(dims, 0), api))
i1 = ctx.to_device(synth.post_process(synth.two_drop_phi(dims, 1), api))
vel = ctx.grid(dims, 4)
try:
    it = ctx.optical_flow4d(vel, i0, i1, None, 1e-3, 1e-4, 0., 1e-2, -1.)
    print("iterations", it)
except Exception as e:  # noqa: BLE001
    print("ERROR", e)
st = ctx.seq_stats()
n = max(st["dots"], 1)
print(st)
print("per dot: dirty leaves %.1f  raw %.1f  pieces %.1f  careful segments %.1f" % (
    st["dirty_leaves"] / n, st["raw_products"] / n, st["pieces"] / n, st["careful_segments"] / n))
clk = 1.965e3  # cycles per microsecond at the B200's 1965 MHz
print("resolver per dot: gather %.1f us  compose %.1f us  walk %.1f us (%.0f steps)  finish %.1f us" % (
    st["cyc_gather"] / n / clk, st["cyc_compose"] / n / clk, st["cyc_walk"] / n / clk, st["walk_steps"] / n, st["cyc_finish"] / n / clk))
ctx.close()
