#!/usr/bin/env python
"""bench.py -- FlOF mode-1 4D deformation solve on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--res 64] [--impl ours|reference]

A "step" is one full mode-1 deformation solve (opticalFlowMultiscale4d with the README solver
parameters of scenes/flof.py:304-320, 913-915: pyramid, 3 OF steps per level, final SDF
projection) on a synthetic two-drop 4D SDF pair (BASELINE.json configs[3]: 64^4).

One JSON line:
  value        seconds per solve, inputs resident in HBM (CUDA-event timed on the library's stream)
  e2e          the same solve through the host-buffer plugin call flof_optical_flow_multiscale4d_host:
               H2D of i0, i1, vel from pinned memory and D2H of the deformation inside the timed region
  roofline     dominant kernel: algorithmic bytes per launch / its average duration, measured live with
               CUDA events bracketing every launch (flof_profile_begin/end) in K extra steps
  cpu_baseline the reference's own CPU implementation (oracle/_ref, else our C port) on a bounded sample of the SAME
               64^4 workload: the first finest-level opticalFlow4d solve (assembly + CG), scaled by CG cell-updates
  config.parity   the result of the timed solve against the reference run's golden (tests/golden/mode1_64x64.npz):
               CG stopping iterations, error trace, deformation bits on the stored lattice; at N > 1 additionally
               against the single-GPU record (tests/golden/bench_n1_record.json)
  config.res128   the north-star configuration (128^4, BASELINE.json configs[4]) measured in the same process:
               1 warm-up + 2 solves, value / cg_iters / final_error / dominant kernel
--impl reference times the reference's CPU implementation on the SAME configuration (one real solve, unscaled).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "flof_mode1_4d_deformation_solve_wall_time"
UNIT = "s"

# algorithmic bytes per cell and launch (SURVEY.md §8d / DESIGN.md §4)
KERNEL_BYTES_PER_CELL = {
    "k_cg_apply": 48,         # srch 16 + grad 16 -> tmp 16
    "k_cg_update": 128,       # result, res, srch, tmp, grad 80 -> result, res, res*precond 48
    "k_cg_direction": 48,     # res*precond, srch 32 -> srch 16
    "k_dot_seq": 32,          # the two vectors of a sequential-order dot product, read once
    "k_gauss_blur4d": 32,     # Vec4 in 16 -> Vec4 out 16 per pass
    "k_cv_expol_blur4d": 36,  # a 16 + marker 4 -> tmp 16 (dense kernel: every cell read and written)
    "k_cv_expol_planes": 36,  # component-plane work-list kernel (default): same algorithmic sweep
    "k_cv_expol_items": 36,   # work-list kernel: same algorithmic sweep (only marker == 0 cells are recomputed)
    "k_project_cells": 44,    # vel 16, phiOrg 4, phiTarget 4 -> dst 16, marker 4
    "k_semi_lagrange4d<float4>": 48,
    "k_semi_lagrange4d<float>": 24,
    "k_of_assemble": 40,
}


def peaks():
    fn = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(fn):
        try:
            return float(json.load(open(fn))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_reference_solve(res, threads=None):
    """One mode-1 solve by the reference's CPU implementation on a `res`^4 synthetic pair.
    Returns (seconds, kind, cores, cg_iters)."""
    from oracle import ref
    from ofblend_b200 import synth
    if ref.available():
        mod, kind = ref, "reference"
    else:
        from oracle import port as mod
        kind = "port"
    cores = mod.set_threads(threads or (os.cpu_count() or 1))

    class Ops:
        set_bound4d = staticmethod(lambda a, v, w: mod.set_bound4d(a, v, w))
        extrap4d_ls_simple = staticmethod(lambda a, d, i: mod.extrap4d_ls_simple(a, d, i))
        mult_const = staticmethod(lambda a, s: mod.grid_op4d("multConst", a, None, s))

    dims = (res, res, res, res)
    i0 = synth.post_process(synth.two_drop_phi(dims, 0), Ops)
    i1 = synth.post_process(synth.two_drop_phi(dims, 1), Ops)
    v0 = np.zeros(i0.shape + (4,), np.float32)
    t0 = time.time()
    mod.optical_flow_multiscale4d(v0, i0, i1, **synth.MODE1_PARAMS)
    return time.time() - t0, kind, cores


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation (all host threads) on the benchmarked configuration.
    One REAL solve at `--res` (64^4: several minutes), no scaling; --steps/--warmup are not honoured beyond that -- a
    CPU solve has no clocks to settle and K = 20 of them would take hours -- and the line says so (steps 1, warmup 0).
    Only 128^4 (hours, > 90 GB) is extrapolated from the measured 64^4 solve by cell count and labelled as such."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample_res = min(args.res, 64)
    scale = (args.res / sample_res) ** 4
    sec, kind, cores = cpu_reference_solve(sample_res)
    if scale == 1:
        sample = ("%s CPU implementation (OpenMP build, %d threads), ONE full mode-1 solve on the %d^4 synthetic two-drop pair: "
                  "the benchmarked configuration itself, measured, not scaled" % (kind, cores, args.res))
    else:
        sample = ("%s CPU implementation (%d threads), one full mode-1 solve on the %d^4 synthetic pair measured (%.1f s), "
                  "EXTRAPOLATED x%d by cell count to %d^4" % (kind, cores, sample_res, sec, int(scale), args.res))
    sec *= scale
    line = {"metric": METRIC, "value": sec, "unit": UNIT, "n_gpus": args.gpus, "steps": 1, "warmup": 0,
            "requested": {"steps": args.steps, "warmup": args.warmup},
            "ms_per_step": sec * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": workload_config(args.res, 1, reference=True),
            "cpu_baseline": {"value": sec, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": sec, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def workload_config(res, world, reference=False):
    cells = res ** 4
    cfg = {"workload": "flof mode 1 (opticalFlowMultiscale4d, README parameters: wSmooth 1e-3, wEnergy 1e-4, "
                       "cgAccuracy 1e-2, postVelBlur 4, multiStep 3, minGridSize 20, final projection) on "
                       "synthetic two-drop 4D SDF pair %d^4%s" % (res, {64: " (BASELINE.json configs[3])", 128: " (BASELINE.json configs[4])"}.get(res, "")),
           "res": res, "levels": [res >> l for l in range(8) if (res >> l) > 10 and (l == 0 or (res >> (l - 1)) > 20)]}
    if not reference:
        cfg["l2"] = "working set (%d MB of grids) exceeds the 126 MB L2; no explicit flush" % (cells * 4 * 30 // 2 ** 20)
        cfg["parallelism"] = ("1 GPU" if world == 1 else
                              "t-sharded over %d GPUs (levels >= 2^22 cells; halos and CG scalars through NVLink peer "
                              "mailboxes, NCCL all-gathers)" % world)
    return cfg


def bits_checksum(a):
    """order-independent integer checksum of a float32 array's bit patterns (exactly reproducible, unlike an fp sum)"""
    u = np.ascontiguousarray(a).view(np.uint32).ravel()
    return "%016x-%08x" % (int(u.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(u)))


N1_RECORD = os.path.join(ROOT, "tests", "golden", "bench_n1_record.json")


def parity_record(res, vel_h, cg_iters, errs, world):
    """What the timed solve produced, against the reference run's golden (64^4) and the committed single-GPU record."""
    out = {"cg_iters": cg_iters, "final_error": errs[-1] if errs else None, "deformation_checksum": bits_checksum(vel_h)}
    gfn = os.path.join(ROOT, "tests", "golden", "mode1_%dx%d.npz" % (res, res))
    if os.path.isfile(gfn):
        g = np.load(gfn)
        st = int(g["stride"])
        sub = vel_h.reshape(res, res, res, res, 4)[::st, ::st, ::st, ::st]
        d = sub.astype(np.float64) - g["vel_sub"]
        out["vs_reference_run"] = {
            "golden": "tests/golden/mode1_%dx%d.npz (unmodified reference, tests/golden/make_golden.py)" % (res, res),
            "cg_iters_equal": cg_iters == [int(x) for x in g["cg_iters"]],
            "error_trace_equal_1e-6": bool(len(errs) == len(g["errs"]) and np.allclose(errs, g["errs"], rtol=1e-6)),
            "deformation_bit_identical_on_lattice": bool(np.array_equal(sub, g["vel_sub"])),
            "deformation_rel_l2": float(np.linalg.norm(d.ravel()) / max(np.linalg.norm(g["vel_sub"].astype(np.float64).ravel()), 1e-300)),
            "deformation_max_abs_cells": float(np.abs(d).max())}
    try:
        rec = json.load(open(N1_RECORD)).get(str(res))
    except Exception:
        rec = None
    if world < 0:
        pass  # (a run in another dot mode: not comparable with the stored record)
    elif rec is not None:
        out["parity_vs_n1"] = bool(rec["cg_iters"] == cg_iters and rec["deformation_checksum"] == out["deformation_checksum"])
        out["n1_record"] = "tests/golden/bench_n1_record.json (single-GPU run of this bench)"
    elif world == 1:
        out["parity_vs_n1"] = True
    return out


def measure(ctx, api, fdist, res, steps, warmup, world, local_rank, with_e2e, prof_steps, check_n1=True):
    """Times `steps` mode-1 solves at res^4 (after `warmup`); returns a dict of everything the JSON line needs."""
    from ofblend_b200 import capi, synth
    dims = (res, res, res, res)
    cells = res ** 4
    # inputs: synthetic two-drop pair, pre-processed on the GPU by the product kernels (outside the timed region)
    i0_h = synth.post_process(synth.two_drop_phi(dims, 0), api)
    i1_h = synth.post_process(synth.two_drop_phi(dims, 1), api)
    i0 = ctx.to_device(i0_h)
    i1 = ctx.to_device(i1_h)
    vel = ctx.grid(dims, 4)
    params = capi.make_params(**synth.MODE1_PARAMS)
    zero4 = np.zeros(4, np.float32)

    def barrier():
        fdist.barrier(ctx)  # stream synchronise + NCCL barrier over all ranks

    def one_step():
        ctx.grid_set_const(vel, zero4)
        return ctx.optical_flow_multiscale4d(vel, i0, i1, params, want_trace=True)

    for _ in range(warmup):
        one_step()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launches
    barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    cg_ms = 0.0
    cg_updates = 0
    trace = None
    for _ in range(steps):
        err, trace = one_step()
        dev_ms += trace.total_ms
        n = min(trace.n_solves, 64)
        cg_ms += sum(trace.cg_ms[:n])
        cg_updates += sum(int(trace.cg_iters[q]) * int(trace.cg_cells[q]) for q in range(n))
    barrier()
    wall = time.perf_counter() - t0
    launches = ctx.launches - launches0
    out = {"res": res, "cells": cells}
    # device time (CUDA events on the library's stream, recorded inside the call); max over ranks
    out["ms_per_step"] = fdist.max_over_ranks(ctx, dev_ms / steps)
    out["wall_ms_per_step"] = wall / steps * 1e3
    out["launches"] = int(launches)
    out["cg_cell_updates_per_s"] = cg_updates / max(cg_ms * 1e-3, 1e-12)
    out["cg_updates_per_step"] = cg_updates / steps
    out["cg_iters"] = [int(trace.cg_iters[q]) for q in range(min(trace.n_solves, 64))]
    out["cg_cells"] = [int(trace.cg_cells[q]) for q in range(min(trace.n_solves, 64))]
    out["errs"] = [float(trace.errs[q]) for q in range(min(trace.n_errs, 64))]
    out["seq"] = ctx.seq_stats()
    # what the timed solve produced, checked on the host (outside every timed region)
    out["parity"] = parity_record(res, vel.download(), out["cg_iters"], out["errs"], world if check_n1 else -1)

    # ---- e2e: host buffers through the plugin-level C-ABI call, copies inside the timed region
    if with_e2e:
        hb = {}
        for name, nbytes in (("i0", cells * 4), ("i1", cells * 4), ("vel", cells * 16)):
            p = C.c_void_p()
            ctx._chk(ctx.lib.flof_host_alloc(ctx.h, C.byref(p), C.c_size_t(nbytes)))
            hb[name] = (p, np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(nbytes // 4,)))
        hb["i0"][1][:] = i0_h.ravel()
        hb["i1"][1][:] = i1_h.ravel()
        e2e_times = []
        if world > 1:
            ctx.set_option("host_result_rank", 0)   # the job's result is read once, by rank 0 (the rank that would write the file)
        tr2 = capi.MultiscaleTrace()
        e2 = C.c_float(0)
        for s in range(1 + steps):
            hb["vel"][1][:] = 0.0
            barrier()
            t1 = time.perf_counter()
            ctx._chk(ctx.lib.flof_optical_flow_multiscale4d_host(ctx.h, hb["vel"][0], hb["i0"][0], hb["i1"][0],
                                                                 capi.Dim4(*dims), C.byref(params), C.byref(tr2), C.byref(e2)))
            float(hb["vel"][1][:4096].sum())  # touch the result on the host
            barrier()
            if s > 0:
                e2e_times.append(time.perf_counter() - t1)
        out["e2e_s"] = fdist.max_over_ranks(ctx, float(np.mean(e2e_times)))
        # the e2e call must deliver the same bits as the resident call (it is the same solve behind host copies)
        if int(os.environ.get("RANK", "0")) == 0:
            out["parity"]["e2e_result_equals_resident"] = bool(bits_checksum(hb["vel"][1]) == out["parity"]["deformation_checksum"])
        if world > 1:
            ctx.set_option("host_result_rank", -1)
        for name in hb:
            ctx.lib.flof_host_free(ctx.h, hb[name][0])
    sampler.stop_flag = True
    sampler.join(timeout=2)
    out["clocks"] = sampler.summary()

    # ---- live per-kernel timing: CUDA events bracket every launch of K more steps of the same workload
    # (kept out of the steps timed above so the ~2 event records per launch do not perturb `value`)
    stats = (KernelStat * 256)()
    nstat = C.c_int(0)
    ctx._chk(ctx.lib.flof_profile_begin(ctx.h))
    for _ in range(prof_steps):
        one_step()
    ctx._chk(ctx.lib.flof_profile_end(ctx.h, stats, 256, C.byref(nstat)))
    tot_ms = sum(stats[q].total_ms for q in range(nstat.value))
    peak, peak_src = peaks()
    ktab = []
    for q in range(nstat.value):
        nm = stats[q].name.decode().replace("(", "").replace(")", "")
        key = next((k for k in KERNEL_BYTES_PER_CELL if nm.startswith(k)), None)
        row = {"kernel": nm, "cells": int(stats[q].cells), "launches_per_step": stats[q].launches / prof_steps,
               "ms_per_step": stats[q].total_ms / prof_steps, "avg_launch_ms": stats[q].total_ms / stats[q].launches,
               "share": stats[q].total_ms / max(tot_ms, 1e-9)}
        if key is not None and stats[q].cells > 0:
            # a sharded level is cut along t: each rank's launch covers cells / world of the level
            sharded = world > 1 and int(stats[q].cells) >= (1 << 22)
            row["cells_per_launch"] = int(stats[q].cells) // (world if sharded else 1)
            row["bytes_per_launch"] = KERNEL_BYTES_PER_CELL[key] * row["cells_per_launch"]
            row["gbs"] = row["bytes_per_launch"] / (row["avg_launch_ms"] * 1e-3) / 1e9
            row["frac"] = row["gbs"] / peak
        ktab.append(row)
    out["kernels"] = ktab
    out["peak"], out["peak_src"] = peak, peak_src
    out["prof_steps"] = prof_steps
    # step-level roofline: algorithmic bytes of all classified launches / step time / peak
    alg = sum(r["bytes_per_launch"] * r["launches_per_step"] for r in ktab if "bytes_per_launch" in r)
    out["step_roofline_frac"] = alg / (out["ms_per_step"] * 1e-3) / 1e9 / peak
    for g in (i0, i1, vel):
        g.free()
    return out


def roofline_of(m):
    ktab = m["kernels"]
    dom = next((r for r in ktab if "gbs" in r), None)  # ktab is sorted by total time: dominant (kernel, level)
    if dom is None:
        return {"bound": "hbm", "achieved": None, "peak": m["peak"], "unit": "GB/s", "frac": None, "traffic": None}
    # measured DRAM traffic of that kernel: one `ncu --set full` capture per round (profiles/*_ncu_traffic.json),
    # per cell, scaled to the cells of this launch
    traffic, traffic_src = None, None
    for fn in ("r2_ncu_traffic.json", "r1_ncu_traffic.json"):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", fn)))
            tk = next((k for k in tj["kernels"] if dom["kernel"].startswith(k)), None)
            if tk is not None:
                traffic = tj["kernels"][tk]["bytes_per_cell"] * dom["cells_per_launch"]
                traffic_src = "profiles/%s (ncu --set full at 64^4, %.1f B/cell%s)" % (
                    fn, tj["kernels"][tk]["bytes_per_cell"], ", captured at commit %s" % tj["commit"] if "commit" in tj else "")
                break
        except Exception:
            pass
    return {"bound": "hbm", "kernel": dom["kernel"], "cells_per_launch": dom["cells_per_launch"], "achieved": dom["gbs"],
            "peak": m["peak"], "unit": "GB/s", "frac": dom["frac"], "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": m["peak_src"], "bytes_per_launch": dom["bytes_per_launch"], "avg_launch_ms": dom["avg_launch_ms"],
            "share_of_step": dom["share"], "step_frac": m["step_roofline_frac"],
            "how": "CUDA events around every launch on the library stream, %d steps" % m["prof_steps"]}


def cpu_baseline_sample(res, cg_iters, cg_cells, api):
    """cpu_baseline: the reference's opticalFlow4d (assembly + Jacobi-PCG, ref optflow4d.cpp:361-553) on the finest level of
    the benchmarked pair, without the blur, timed on the host cores at two CG accuracies.  The difference of the two runs
    gives the cost of a CG iteration, the rest the cost of the assembly; both are single-threaded in the reference whatever
    the build and scale with the cell count, so the assembly + CG share of the whole mode-1 solve follows from the product's
    own list of solves (cells and iterations per solve; bit-identical CG: same stopping iterations as the reference)."""
    from oracle import ref
    from ofblend_b200 import synth
    if ref.available():
        mod, kind = ref, "reference"
    else:
        from oracle import port as mod
        kind = "port"
    cores = mod.set_threads(os.cpu_count() or 1)
    dims = (res, res, res, res)
    cells = res ** 4
    i0 = synth.post_process(synth.two_drop_phi(dims, 0), api)   # same inputs as the GPU run (bit-identical pre-processing)
    i1 = synth.post_process(synth.two_drop_phi(dims, 1), api)
    v0 = np.zeros(i0.shape + (4,), np.float32)
    runs = []
    for acc in (5e-1, 1e-1):  # loose accuracies bound the sample to ~10 s; the per-iteration cost does not depend on them
        _, it = api.optical_flow4d(v0, i0, i1, 1e-3, 1e-4, 0., acc, 0.1, want_iters=True)   # iteration count from the product
        t0 = time.time()
        mod.optical_flow4d(v0, i0, i1, 1e-3, 1e-4, 0., acc, 0.1)
        runs.append((int(it), time.time() - t0))
    (it1, t1), (it2, t2) = runs
    per_iter = (t2 - t1) / (it2 - it1) if it2 > it1 else t2 / max(it2, 1)
    setup = max(t1 - it1 * per_iter, 0.)
    est = sum(per_iter * it * c / cells + setup * c / cells for it, c in zip(cg_iters, cg_cells))
    return {"value": est, "unit": UNIT, "cores": cores, "kind": kind,
            "sample_seconds": t1 + t2, "sample_cg_iterations": [it1, it2], "cpu_cg_cell_updates_per_s": cells / per_iter,
            "cpu_seconds_per_cg_iteration": per_iter, "cpu_seconds_assembly": setup,
            "sample": "%s CPU implementation: opticalFlow4d(wSmooth 1e-3, wEnergy 1e-4, postVelBlur 0) on the %d^4 level of the "
                      "benchmarked pair at cgAccuracy 5e-1 / 1e-1 = assembly + %d / %d CG iterations, %.1f s + %.1f s measured "
                      "-> %.2f s per CG iteration, %.1f s per assembly; `value` = the assembly + CG share of the whole mode-1 "
                      "solve (%d solves, cells and iterations per solve from the timed run), blur / projection / advection of "
                      "the CPU path NOT included -- the reference arm (--impl reference) measures the whole solve"
                      % (kind, res, it1, it2, t1, t2, per_iter, setup, len(cg_iters))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--res", type=int, default=64)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-res128", action="store_true", help="skip the 128^4 sub-record (config.res128)")
    ap.add_argument("--no-tree-record", action="store_true", help="skip the dot_mode 0 sub-record (config.tree_dot_mode)")
    ap.add_argument("--write-n1-record", action="store_true", help="store this run's result as the single-GPU parity record")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    # rank 0 prints exactly one JSON line on stdout: NCCL writes its version banner / debug lines to fd 1, so
    # stdout is pointed at stderr until the line is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    from ofblend_b200 import capi
    from ofblend_b200 import dist as fdist
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # one process per GPU; the NCCL communicator lives inside libflof_b200.so (ofblend_b200/dist.py)
    ctx, rank, world = fdist.init()
    api = capi.HostAPI(ctx)
    res = args.res
    cells = res ** 4

    m = measure(ctx, api, fdist, res, args.steps, args.warmup, world, local_rank, True, max(1, min(args.steps, 3)))
    config = workload_config(res, world)
    config["parity"] = m["parity"]
    config["parity_vs_n1"] = m["parity"].get("parity_vs_n1")
    config["seq_dot_stats"] = m["seq"]

    if res == 64 and not args.no_res128:
        # the north-star configuration in the same process (1 warm-up + 2 solves; BASELINE.json configs[4])
        try:
            m2 = measure(ctx, api, fdist, 128, 2, 1, world, local_rank, False, 1)
            dom2 = roofline_of(m2)
            config["res128"] = {"value": m2["ms_per_step"] / 1e3, "unit": UNIT, "steps": 2, "warmup": 1,
                                "cg_iters": m2["cg_iters"], "final_error": m2["errs"][-1] if m2["errs"] else None,
                                "cg_cell_updates_per_s": m2["cg_cell_updates_per_s"],
                                "deformation_checksum": m2["parity"]["deformation_checksum"],
                                "parity_vs_n1": m2["parity"].get("parity_vs_n1"),
                                "dominant_kernel": {k: dom2.get(k) for k in ("kernel", "frac", "avg_launch_ms", "share_of_step")},
                                "step_roofline_frac": m2["step_roofline_frac"],
                                "kernels": [{k: r.get(k) for k in ("kernel", "cells", "launches_per_step", "avg_launch_ms", "share", "frac")}
                                            for r in m2["kernels"][:10]]}
        except Exception as e:  # noqa: BLE001
            config["res128"] = {"error": str(e)[:300]}
            m2 = None
    else:
        m2 = None

    if res == 64 and not args.no_tree_record:
        # the same solve with tree-reduced dot products (dot_mode 0, the round-1 arithmetic): faster, not the reference's bits
        try:
            ctx.set_option("dot_mode", 0)
            m0 = measure(ctx, api, fdist, res, 3, 1, world, local_rank, False, 1, check_n1=False)
            config["tree_dot_mode"] = {"note": "dot_mode 0: CG dot products as tree reductions instead of the reference's sequential order "
                                               "(default dot_mode 1, which `value` measures); same kernels otherwise",
                                       "value": m0["ms_per_step"] / 1e3, "unit": UNIT, "steps": 3, "warmup": 1, "cg_iters": m0["cg_iters"],
                                       "vs_reference_run": m0["parity"].get("vs_reference_run")}
        except Exception as e:  # noqa: BLE001
            config["tree_dot_mode"] = {"error": str(e)[:300]}
        finally:
            ctx.set_option("dot_mode", 1)

    if res == 64 and not args.no_tree_record:
        # the same solve with the opt-in separable Gaussian blur (blur_mode 1; exact dot products kept): what the bit-exact
        # (2S+1)^4-tap blur costs, and how far the result moves without it
        try:
            ctx.set_option("blur_mode", 1)
            ms = measure(ctx, api, fdist, res, 3, 1, world, local_rank, False, 1, check_n1=False)
            config["separable_blur_mode"] = {"note": "blur_mode 1 (NOT the default): Gaussian blur as four 1D fp32 passes instead of the "
                                                     "reference's (2S+1)^4-tap sums; everything else as in `value`",
                                             "value": ms["ms_per_step"] / 1e3, "unit": UNIT, "steps": 3, "warmup": 1,
                                             "cg_iters": ms["cg_iters"], "vs_reference_run": ms["parity"].get("vs_reference_run")}
        except Exception as e:  # noqa: BLE001
            config["separable_blur_mode"] = {"error": str(e)[:300]}
        finally:
            ctx.set_option("blur_mode", 0)

    line = {"metric": METRIC, "value": m["ms_per_step"] / 1e3, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "wall_ms_per_step": m["wall_ms_per_step"],
            "cg_cell_updates_per_s": m["cg_cell_updates_per_s"],
            "cg_iters": m["cg_iters"],
            "final_error": m["errs"][-1] if m["errs"] else None,
            "gpu_launches": m["launches"],
            # every rank uploads its own t-slab of the inputs (all-gathered over NVLink) and downloads the complete
            # deformation: whole-job bytes
            "e2e": {"value": m["e2e_s"], "unit": UNIT, "h2d_bytes_per_step": cells * 24 * (1 if res % world == 0 else world),
                    "d2h_bytes_per_step": cells * 16,
                    "bytes_per_rank": {"h2d": cells * 24 // (world if res % world == 0 else 1), "d2h_rank0": cells * 16, "d2h_other_ranks": 0},
                    "api": "flof_optical_flow_multiscale4d_host (pinned host buffers)" + (
                        "; every rank uploads its own t-slab, rank 0 downloads the deformation (option host_result_rank = 0)" if world > 1 else "")},
            "clocks": m["clocks"],
            "roofline": roofline_of(m),
            "kernels": m["kernels"][:16]}
    if rank == 0:
        if args.write_n1_record and world == 1:
            try:
                rec = json.load(open(N1_RECORD))
            except Exception:
                rec = {}
            for mm in (m, m2):
                if mm is not None:
                    rec[str(mm["res"])] = {"cg_iters": mm["cg_iters"], "final_error": mm["errs"][-1],
                                           "deformation_checksum": mm["parity"]["deformation_checksum"]}
            json.dump(rec, open(N1_RECORD, "w"), indent=1)
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_sample(res if res <= 64 else 64, m["cg_iters"], m["cg_cells"], api)
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    barrier_all = fdist.barrier
    barrier_all(ctx)
    ctx.close()
    return 0


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("cells", C.c_int64), ("launches", C.c_int), ("total_ms", C.c_float)]


if __name__ == "__main__":
    sys.exit(main())
