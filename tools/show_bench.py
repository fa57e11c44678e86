"""Pretty-prints a bench.py JSON line: python tools/show_bench.py gpurun_out/bench.json"""
import json
import signal
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)  # `| head` closes the pipe early: exit quietly

d = json.load(open(sys.argv[1]))
print("ms_per_step %.2f  e2e %.4f s  iters %s  final_err %s  launches %s" % (
    d["ms_per_step"], d["e2e"]["value"], d.get("cg_iters"), d.get("final_error"), d.get("gpu_launches")))
print("cg cell-updates/s %.3g   clocks %s" % (d.get("cg_cell_updates_per_s", 0), d.get("clocks")))
r = d.get("roofline", {})
print("roofline:", {k: r.get(k) for k in ("kernel", "achieved", "peak", "frac", "share_of_step")})
for k in d.get("kernels", []):
    print("%-40s cells %9d n %6.1f ms %8.3f avg %7.4f share %5.3f frac %s" % (
        k["kernel"][:40], k["cells"], k["launches_per_step"], k["ms_per_step"], k["avg_launch_ms"], k["share"],
        ("%.3f" % k["frac"]) if "frac" in k else "-"))
if "cpu_baseline" in d:
    print("cpu_baseline:", d["cpu_baseline"])
