"""Build libmg_b200.so (C-ABI shared library: C host drivers + sm_100a CUDA kernels) in-tree.

    python -m pde_multigrid_b200.build        # or: from pde_multigrid_b200.build import build; build()

nvcc cross-compiles for sm_100a without a GPU.  The .so is git-ignored but travels to the GPU box.
Flags: -gencode arch=compute_100a,code=sm_100a -lineinfo -O3; host code with -ffp-contract=off so the
per-level coefficients are computed without FMA contraction, like the reference's host code.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libmg_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.c")))


def up_to_date():
    if not os.path.exists(SO):
        return False
    deps = sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps)


SO_DEBUG = os.path.join(HERE, "libmg_b200_dbg.so")


def build_debug_bounds(verbose=False):
    """libmg_b200_dbg.so: the same library with -DMG_DEBUG_BOUNDS (index asserts in the 3D kernels, csrc/mg3d_device.cuh)."""
    return build(force=True, verbose=verbose, extra=("-DMG_DEBUG_BOUNDS",), so=SO_DEBUG, objdir=os.path.join(HERE, "build", "dbg"))


def build(force=False, verbose=False, extra=(), so=SO, objdir=None):
    if so == SO and not force and up_to_date():
        return SO
    objs = []
    objdir = objdir or os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    common = ["-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler",
              "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall", "-I", os.path.join(HERE, "..", "include"), "-I", CSRC]
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        cmd = [NVCC] + common + list(extra) + ["-c", src, "-o", obj]
        if src.endswith(".cu"):
            cmd += ["-std=c++17", "--expt-relaxed-constexpr", "-Xptxas", "-v" if verbose else "-O3"]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s:\n%s\n" % (src, out))
        elif verbose and out.strip():
            print(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    cmd = [NVCC, "-shared", "-o", so] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    subprocess.run(cmd, check=True)
    return so


if __name__ == "__main__":
    if "--debug-bounds" in sys.argv:
        print(build_debug_bounds(verbose="-v" in sys.argv))
    else:
        print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
