"""Sequential-order dot-product statistics of a full mode-1 solve:  python tools/seq_probe_ms.py RES"""
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
params = capi.make_params(**synth.MODE1_PARAMS)
try:
    err, trace = ctx.optical_flow_multiscale4d(vel, i0, i1, params, want_trace=True)
    print("iterations", [int(trace.cg_iters[q]) for q in range(trace.n_solves)], "ms", trace.total_ms)
except Exception as e:  # noqa: BLE001
    print("ERROR", e)
st = ctx.seq_stats()
n = max(st["dots"], 1)
print(st, "why 0x%x" % st["why"])
print("per dot: dirty leaves %.1f  raw %.1f  pieces %.1f  careful segments %.1f" % (
    st["dirty_leaves"] / n, st["raw_products"] / n, st["pieces"] / n, st["careful_segments"] / n))
ctx.close()
