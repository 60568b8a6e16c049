#!/usr/bin/env python3
"""oracle/build_ref.py -- TEST INFRASTRUCTURE, not product code.

Builds oracle/_ref/libmg_ref.so: the UNMODIFIED reference CPU solver (NOCUDA_TESI, the
parity oracle named by SURVEY.md section 8c) compiled with g++ from the sources where they
lie under /root/reference.  Nothing is copied into the repo; the only outputs are object
files and the shared library under oracle/_ref/ (git-ignored, but shipped to the GPU box).

Variants (one object each, every one in its own C++ namespace):
    ref3d_f32  ref3d_f64  ref3d_f32c  ref3d_f64c     (c = CORRECTED residual signs)
    ref2d_f32  ref2d_f64
    ref1d_f32  ref1d_f64  ref1d_f32c  ref1d_f64c
    ref3d_f32x  ref3d_f64x  ref3d_f32cx  ref3d_f64cx   the 3D variants with -DNDEBUG: assertions compiled out, sources
                     untouched -- the reference then runs non-cubic grids (only N3/Grid3D.cpp:10-11 forbids them)
    ref3d_f64cO0   = ref3d_f64c built with -O0, i.e. as shipped (the reference's CompileAndLink passes no flags);
                     only bench.py's CPU baseline times it, next to the -O2 figure

The CORRECTED variants include a patched copy of MultiGrid3D.cpp / MultiGrid1D.cpp that is
written to a temporary directory at build time and deleted afterwards.  Each patch must
change exactly one line or the build fails:
    N3/MultiGrid3D.cpp:723   ((N-2*v[idx]-S)/h_y2) -> +S ,  ((D-2*v[idx]-U)/h_z2) -> +U
    N1/MultiGrid1D.cpp:210   - h_v[posX]/(exp(xj)+1) -> + h_v[posX]/(exp(xj)+1)

Flags: -O2, never -ffast-math (changes bits), no -march (keeps x86-64 baseline: no FMA
contraction, like the reference's own CompileAndLink which passes no flags at all).
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("MG_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

DIRS = {
    "3d": os.path.join(REF_ROOT, "NOCUDA_TESI", "POISSON_3D(TESI)"),
    "2d": os.path.join(REF_ROOT, "NOCUDA_TESI", "PDE Lyapunov 2D"),
    "1d": os.path.join(REF_ROOT, "NOCUDA_TESI", "EQUAZIONE 1D"),
}

PATCHES = {
    "3d": ("MultiGrid3D.cpp", [
        ("((N-2*v[idx]-S)/h_y2)", "((N-2*v[idx]+S)/h_y2)"),
        ("((D-2*v[idx]-U)/h_z2)", "((D-2*v[idx]+U)/h_z2)"),
    ]),
    "1d": ("MultiGrid1D.cpp", [
        ("- h_v[posX]/(exp(xj)+1);", "+ h_v[posX]/(exp(xj)+1);"),
    ]),
}


def reference_available():
    return all(os.path.isdir(d) for d in DIRS.values())


def _patched_dir(dim, tmp):
    fname, subs = PATCHES[dim]
    with open(os.path.join(DIRS[dim], fname), "r", encoding="latin-1") as fh:
        src = fh.read()
    lines_before = src.split("\n")
    for old, new in subs:
        if src.count(old) != 1:
            raise RuntimeError("patch target %r occurs %d times in %s (expected 1)" % (old, src.count(old), fname))
        src = src.replace(old, new)
    lines_after = src.split("\n")
    changed = sum(1 for a, b in zip(lines_before, lines_after) if a != b)
    if changed != 1 or len(lines_before) != len(lines_after):
        raise RuntimeError("patch of %s changed %d lines (expected exactly 1)" % (fname, changed))
    d = os.path.join(tmp, "patched_" + dim)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, fname), "w", encoding="latin-1") as fh:
        fh.write(src)
    return d


def build(verbose=False):
    if not reference_available():
        raise RuntimeError("reference sources not found under %s" % REF_ROOT)
    os.makedirs(OUT, exist_ok=True)
    cxx = os.environ.get("CXX", "g++")
    common = [cxx, "-O2", "-fPIC", "-w", "-std=gnu++17", "-fno-fast-math", "-ffp-contract=off", "-c"]
    objs = []
    tmp = tempfile.mkdtemp(prefix="mg_ref_build_")
    try:
        variants = []
        for dim in ("3d", "2d", "1d"):
            for prec in ("f32", "f64"):
                variants.append((dim, prec, False))
                if dim in PATCHES:
                    variants.append((dim, prec, True))
        variants = [v + ("O2", False) for v in variants] + [("3d", "f64", True, "O0", False)]
        variants += [("3d", prec, corr, "O2", True) for prec in ("f32", "f64") for corr in (False, True)]
        for dim, prec, corrected, opt, ndebug in variants:
            prefix = "ref%s_%s%s%s%s" % (dim, prec, "c" if corrected else "", "" if opt == "O2" else opt, "x" if ndebug else "")
            obj = os.path.join(OUT, prefix + ".o")
            cmd = [("-" + opt) if a == "-O2" else a for a in common]
            cmd += ["-DREF_PREFIX=" + prefix]
            if prec == "f64":
                cmd += ["-DREF_F64"]
            if ndebug:
                cmd += ["-DNDEBUG"]
            if corrected:
                cmd += ["-I", _patched_dir(dim, tmp)]
            cmd += ["-I", DIRS[dim], "-I", HERE, os.path.join(HERE, "ref_wrap%s.cpp" % dim), "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.run(cmd, check=True)
            objs.append(obj)
        so = os.path.join(OUT, "libmg_ref.so")
        subprocess.run([cxx, "-shared", "-o", so] + objs + ["-lm"], check=True)
        for o in objs:
            os.remove(o)
        return so
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
