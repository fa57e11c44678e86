"""One mode-1 step for ncu: `ncu --profile-from-start off ... python tools/profile_step.py [res]`.
Pre-processing and one warm-up solve run outside the cudaProfilerStart/Stop window."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ofblend_b200 import capi, synth  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dims = (res, res, res, res)
ctx = capi.Context(0)
api = capi.HostAPI(ctx)
i0 = ctx.to_device(synth.post_process(synth.two_drop_phi(dims, 0), api))
i1 = ctx.to_device(synth.post_process(synth.two_drop_phi(dims, 1), api))
vel = ctx.grid(dims, 4)
params = capi.make_params(**synth.MODE1_PARAMS)
ctx.optical_flow_multiscale4d(vel, i0, i1, params)
cudart = None
for name in ("libcudart.so.12", "libcudart.so"):
    try:
        cudart = ctypes.CDLL(name)
        break
    except OSError:
        pass
if cudart:
    cudart.cudaProfilerStart()
for _ in range(steps):
    ctx.grid_set_const(vel, np.zeros(4, np.float32))
    err, tr = ctx.optical_flow_multiscale4d(vel, i0, i1, params, want_trace=True)
ctx.sync()
if cudart:
    cudart.cudaProfilerStop()
print("step done: %.2f ms, final error %.6g, launches %d" % (tr.total_ms, err, ctx.launches))
