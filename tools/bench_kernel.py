"""Micro-benchmark of single operators at a given resolution (live CUDA-event timing through
flof_profile_begin/end):  python tools/bench_kernel.py 64 expol gauss cg advect project"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ofblend_b200 import capi, synth  # noqa: E402


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("cells", C.c_int64), ("launches", C.c_int), ("total_ms", C.c_float)]


def main():
    res = int(sys.argv[1])
    which = sys.argv[2:] or ["expol", "gauss", "cg", "advect", "project"]
    dims = (res, res, res, res)
    cells = res ** 4
    ctx = capi.Context(0)
    api = capi.HostAPI(ctx)
    i0 = ctx.to_device(synth.post_process(synth.two_drop_phi(dims, 0), api))
    i1 = ctx.to_device(synth.post_process(synth.two_drop_phi(dims, 1), api))
    vel = ctx.grid(dims, 4)
    dst = ctx.grid(dims, 4)
    mk = ctx.grid(dims, 1)
    ctx.optical_flow4d(vel, i0, i1, None, 1e-3, 1e-4, 4., 1e-2, 0.1)      # a realistic deformation
    ctx.project_cells(dst, vel, i0, i1, mk, 4., 40)                       # realistic marker / projected field

    def run(name):
        if "@" in name:
            name, v = name.split("@")          # "expol@7.1": expol_mode 7, expol_variant 1
            mode, _, var = v.partition(".")
            ctx.set_option("expol_mode", int(mode))
            ctx.set_option("expol_variant", int(var or 0))
        if name == "expol":
            ctx.cv_expol_blur4d(dst, mk, 8)
        elif name == "gauss":
            ctx.gaussian_blur4d(vel, 2.0, 1)
        elif name == "gauss1":
            ctx.gaussian_blur4d(vel, 1.0, 1)
        elif name == "cg":
            v = ctx.grid(dims, 4)
            ctx.optical_flow4d(v, i0, i1, None, 1e-3, 1e-4, 0., 1e-2, -1.)
            v.free()
        elif name == "advect":
            ctx.advect_cfl4d(999., vel, dst)
            g = ctx.grid(dims, 1)
            ctx.advect_cfl4d(999., vel, g)
            g.free()
        elif name == "project":
            ctx.project_cells(dst, vel, i0, i1, mk, 4., 40)

    for name in which:
        run(name)
        stats = (KernelStat * 64)()
        n = C.c_int(0)
        ctx._chk(ctx.lib.flof_profile_begin(ctx.h))
        for _ in range(3):
            run(name)
        ctx._chk(ctx.lib.flof_profile_end(ctx.h, stats, 64, C.byref(n)))
        for q in range(n.value):
            s = stats[q]
            if s.total_ms / s.launches > 0.004:
                print("%-8s %-44s launches %4d  avg %8.4f ms   %7.1f ns/Mcell" % (
                    name, s.name.decode()[:44], s.launches, s.total_ms / s.launches, s.total_ms / s.launches * 1e6 / (cells / 1e6) / 1e3))
    ctx.close()


if __name__ == "__main__":
    main()
