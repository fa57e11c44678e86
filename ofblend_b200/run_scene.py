"""Runs a mantaflow scene script (e.g. the reference's scenes/flof.py) against the B200 `manta`
module -- the replacement for `./manta scenes/flof.py <args>` (ref pwrapper/pymain.cpp:40-127).

    python -m ofblend_b200.run_scene /path/to/scenes/flof.py dataid0 0 dataid1 1 mode 1

* `import manta` resolves to ofblend_b200/host/manta*.so (C++ host layer over the C ABI).
* The scene's directory is put on sys.path like the reference does, so `from ofHelpers import *`
  works.  ofHelpers.py:235 of the reference contains one Python-2 `print` statement (a SyntaxError
  under Python 3, SURVEY.md correction 4); it is fixed on the fly while loading -- the file on disk
  is not modified and no other line is touched.
* Working directory, command-line key/value pairs and output files are exactly the reference's.
"""
import os
import re
import runpy
import sys
import types

HOST_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host")


def _load_helpers_py3(scene_dir):
    path = os.path.join(scene_dir, "ofHelpers.py")
    if not os.path.isfile(path) or "ofHelpers" in sys.modules:
        return
    src = open(path).read()
    fixed = re.sub(r'^(\s*)print\s+"([^"\n]*)"\s*;?\s*$', r'\1print("\2")', src, flags=re.M)
    if fixed == src:
        return  # already valid Python 3: let the normal import machinery handle it
    mod = types.ModuleType("ofHelpers")
    mod.__file__ = path
    sys.modules["ofHelpers"] = mod
    exec(compile(fixed, path, "exec"), mod.__dict__)


def main(argv):
    if len(argv) < 2:
        print(__doc__)
        return 2
    scene = os.path.abspath(argv[1])
    scene_dir = os.path.dirname(scene)
    if HOST_DIR not in sys.path:
        sys.path.insert(0, HOST_DIR)
    if scene_dir not in sys.path:
        sys.path.insert(0, scene_dir)
    import manta  # noqa: F401  (fails loudly if the module is not built or there is no GPU)
    sys.argv = [scene] + list(argv[2:])
    _load_helpers_py3(scene_dir)
    try:
        runpy.run_path(scene, run_name="__main__")
    finally:
        # Grid.save() returns before the background deflate + write has finished (overlapped .uni I/O): wait for it
        # here so that a failed write (disk full, I/O error) raises and the process exits non-zero, like the
        # reference's synchronous save() would have (ref fileio.cpp:834-902)
        manta.flushUniWrites()
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
