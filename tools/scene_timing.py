"""Wall-clock of the unmodified scenes/flof.py on the README data set (BASELINE.json configs[0..2]) through the B200
`manta` module, with the background .uni I/O on and off:  python tools/scene_timing.py"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENE = os.path.join(ROOT, "oracle", "_ref", "scenes", "flof.py")
DATA = os.path.join(ROOT, "oracle", "_ref", "data", "readme")


def run(cwd, env_extra, *args):
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    env.update(env_extra)
    t0 = time.perf_counter()
    p = subprocess.run([sys.executable, "-m", "ofblend_b200.run_scene", SCENE] + [str(a) for a in args], cwd=cwd, env=env,
                       capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        print(p.stdout[-2000:], p.stderr[-2000:])
        raise SystemExit(1)
    return dt


def main():
    # the first pass only warms the page cache of the input files and is not reported
    for label, env in (("(page-cache warm-up)", {}), ("async I/O (default)", {}), ("FLOF_SYNC_IO=1", {"FLOF_SYNC_IO": "1"}),
                       ("async I/O (default)", {})):
        d = tempfile.mkdtemp()
        for f in os.listdir(DATA):
            if not f.startswith("ref_"):
                os.symlink(os.path.join(DATA, f), os.path.join(d, f))
        run(d, env, "dataid0", 0, "dataid1", 1, "mode", 1)  # warm-up of the process start / page cache, also produces defo 0->1
        t1 = run(d, env, "dataid0", 1, "dataid1", 0, "mode", 1)
        t2 = run(d, env, "dataid0", 0, "dataid1", 1, "mode", 2, "writeuni", 1)
        t3 = run(d, env, "mode", 3, "twoway", 1, "alpha", 50, "writeuni", 1)
        print("%-22s mode 1 %.2f s   mode 2 %.2f s   mode 3 (two-way, alpha 50, 119 frames) %.2f s" % (label, t1, t2, t3))


if __name__ == "__main__":
    main()
