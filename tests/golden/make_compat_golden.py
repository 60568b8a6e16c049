#!/usr/bin/env python3
"""tests/golden/make_compat_golden.py -- golden log/diff.txt files of THE REFERENCE'S OWN CLASSES (test infrastructure).

Compiles tests/compat/drv{1,2,3}d.cpp -- the calls of the reference's main()s with PrintDiff() switched on -- together with the
reference's unmodified Grid?D.cpp / MultiGrid?D.cpp straight from /root/reference (g++ -O2, no copy into this repo), runs the
binaries on this CPU and stores what they write to log/diff.txt under tests/golden/compat_diff_?d.txt.  The GPU test
(tests/test_compat_equivalence.py) builds the same drivers against the shim of include/compat/ + libmg_b200.so and demands
the same bytes.  Needs /root/reference:   python tests/golden/make_compat_golden.py
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MG_REFERENCE_ROOT", "/root/reference")
CASES = {  # dim: (reference directory, sources, driver arguments: n v0 nu)
    "3d": (os.path.join(REF, "NOCUDA_TESI", "POISSON_3D(TESI)"), ("Grid3D.cpp", "MultiGrid3D.cpp"), ("17", "2", "3")),
    "2d": (os.path.join(REF, "NOCUDA_TESI", "PDE Lyapunov 2D"), ("Grid2D.cpp", "MultiGrid2D.cpp"), ("33", "1", "20")),
    "1d": (os.path.join(REF, "NOCUDA_TESI", "EQUAZIONE 1D"), ("Grid1D.cpp", "MultiGrid1D.cpp"), ("129", "2", "100")),
    # a NON-CUBIC grid (nx ny nz v0 nu): the reference's classes built with -DNDEBUG -- its asserts at N3/Grid3D.cpp:10-11 are all
    # that forbids one; the driver adds a hand-made cycle through the free-array operators
    "3d_box": (os.path.join(REF, "NOCUDA_TESI", "POISSON_3D(TESI)"), ("Grid3D.cpp", "MultiGrid3D.cpp"), ("33", "17", "9", "2", "3")),
}
EXTRA_FLAGS = {"3d_box": ["-DNDEBUG"]}


def build_reference_side(dim, out):
    refdir, srcs, _ = CASES[dim]
    cmd = ["g++", "-O2", "-w"] + EXTRA_FLAGS.get(dim, []) + ["-include", os.path.join(ROOT, "tests", "compat", "ref_malloc_pad.h"), "-I", refdir,
           os.path.join(ROOT, "tests", "compat", "drv%s.cpp" % dim)] + [os.path.join(refdir, s) for s in srcs] + ["-o", out, "-lm"]
    subprocess.run(cmd, check=True)


def main():
    for dim, (_, _, args) in CASES.items():
        with tempfile.TemporaryDirectory() as tmp:
            exe = os.path.join(tmp, "ref_drv%s" % dim)
            build_reference_side(dim, exe)
            os.mkdir(os.path.join(tmp, "log"))
            out = subprocess.run([exe] + list(args), cwd=tmp, capture_output=True, text=True, check=True).stdout
            assert "finestGridSize: %s" % args[0] in out, out
            dst = os.path.join(HERE, "compat_diff_%s.txt" % dim)
            shutil.copyfile(os.path.join(tmp, "log", "diff.txt"), dst)
            print(dim, args, os.path.getsize(dst), "bytes ->", dst)


if __name__ == "__main__":
    sys.exit(main())
