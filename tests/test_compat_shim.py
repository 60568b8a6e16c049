"""Drop-in proof: the reference's own main() files build UNCHANGED against the
header-compatible shim classes of include/compat/ and libmg_b200.so.

CPU part (needs /root/reference, i.e. this container): each main is compiled through a two-line wrapper
TU that includes the shim headers first (same include guards as the reference headers, which therefore
become no-ops) and then the reference main by absolute path; binaries land in tests/_build/ (git-ignored,
shipped to the GPU box).  GPU part: runs those binaries when present.
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")
REF = os.environ.get("MG_REFERENCE_ROOT", "/root/reference")
MAINS = {
    "3d": (os.path.join(REF, "NOCUDA_TESI", "POISSON_3D(TESI)"), "Poisson3DSolver.cpp", ("Grid3D.h", "MultiGrid3D.h")),
    "2d": (os.path.join(REF, "NOCUDA_TESI", "PDE Lyapunov 2D"), "LyapunovSolver.cpp", ("Grid2D.h", "MultiGrid2D.h")),
    "1d": (os.path.join(REF, "NOCUDA_TESI", "EQUAZIONE 1D"), "Poisson1DSolver.cpp", ("Grid1D.h", "MultiGrid1D.h")),
    # the thesis' GPU programs: same classes (the 2D one with a scalar size and PrintMeanAbsoluteError); main.cu is plain C++
    "cuda3d": (os.path.join(REF, "CUDA_TESI", "CUDA Poisson 3D"), "main.cu", ("Grid3D.h", "MultiGrid3D.h")),
    "cuda2d": (os.path.join(REF, "CUDA_TESI", "CUDA Lyapunov 2D"), "main.cu", ("Grid2D.h", "MultiGrid2D.h")),
    "cuda1d": (os.path.join(REF, "CUDA_TESI", "CUDA 1D"), "main.cu", ("Grid1D.h", "MultiGrid1D.h")),
}
CUDA_INCLUDE = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")  # <math_constants.h> of the twins' inclusion.h


def binary(dim):
    return os.path.join(BUILD, "ref_main_%s" % dim)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not mounted")
@pytest.mark.parametrize("dim", ["1d", "2d", "3d", "cuda1d", "cuda2d", "cuda3d"])
def test_reference_main_builds_unchanged_against_the_shim(mg, dim):
    refdir, main, headers = MAINS[dim]
    os.makedirs(BUILD, exist_ok=True)
    wrapper = os.path.join(BUILD, "wrap_main_%s.cpp" % dim)
    with open(wrapper, "w") as fh:
        for h in headers:
            fh.write('#include "%s"\n' % h)  # include/compat/ (found first through -I order; same guards as the reference)
        fh.write('#include "%s"\n' % os.path.join(refdir, main))
    libdir = os.path.join(ROOT, "pde_multigrid_b200")
    cmd = ["g++", "-O2", "-w", "-I", os.path.join(ROOT, "include", "compat"), "-I", os.path.join(ROOT, "include"),
           "-I", refdir, "-I", CUDA_INCLUDE, wrapper, "-o", binary(dim), "-L", libdir, "-lmg_b200",
           "-Wl,-rpath,$ORIGIN/../../pde_multigrid_b200", "-lm"]
    subprocess.run(cmd, check=True)
    os.remove(wrapper)
    # the shim's classes, not the reference's, must have been compiled in: the binary needs the C ABI
    syms = subprocess.run(["nm", "-D", "--undefined-only", binary(dim)], capture_output=True, text=True).stdout
    d = dim[-2:]
    assert "mg%s_create" % d in syms and "mg%s_fmg" % d in syms


@pytest.mark.gpu
@pytest.mark.parametrize("dim,size", [("1d", 8193), ("2d", 1025), ("3d", 129), ("cuda1d", 8193), ("cuda2d", 65), ("cuda3d", 257)])
def test_reference_main_runs_on_the_gpu(dim, size, tmp_path):
    exe = binary(dim)
    if not os.path.exists(exe):
        pytest.skip("tests/_build/ref_main_%s not built (built by the CPU test in a container with the reference)" % dim)
    (tmp_path / "log").mkdir()
    out = subprocess.run([exe], cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr
    assert "finestGridSize: %d" % size in out.stdout  # N3/Poisson3DSolver.cpp:44 and twins
    if dim == "cuda2d":  # C2/main.cu: FMG(2,500,500) at n = 65 on [0,20]^2 and PrintMeanAbsoluteError: thesis Fig. 4.3 prints 5.32
        mae = float(out.stdout.split("MeanAbsoluteError:")[1].split()[0])
        assert abs(mae - 5.32) < 0.01, out.stdout
    if dim == "2d":  # N2/LyapunovSolver.cpp:44 calls PrintDiff(): log/diff.txt, one line per grid point
        diffs = [float(l.rsplit("diff:", 1)[1]) for l in open(tmp_path / "log" / "diff.txt")]
        assert len(diffs) == size * size
        assert np.isfinite(diffs).all() and np.max(np.abs(diffs)) < 0.05  # O(h) upwind discretisation error on [0,1]^2
