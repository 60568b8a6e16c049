"""Runs one golden case (tests/golden/*.npz, produced by the reference itself) through any object with
the RefMG/PortMG interface, or through the product engine, and yields (name, got, want) pairs."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "ref*.npz")))


def parse_name(path):
    base = os.path.basename(path)[:-4]  # ref3d_f64c
    dim = int(base[3])
    dtype = np.float32 if "f32" in base else np.float64
    corrected = base.endswith("c")
    return dim, dtype, corrected


def run_checker(make, g):
    """make(n, range) -> fresh oracle-like object."""
    n, rng_range = int(g["n"]), tuple(g["range"])
    nu, cycles, fmg = int(g["nu"]), int(g["cycles"]), tuple(int(x) for x in g["fmg"])
    out = {}
    m = make(n, rng_range)
    m.v(0)[...] = g["in_v"]
    m.f(0)[...] = g["in_f"]
    m.relax(0, 2)
    out["relax2_v"] = m.v(0).copy()
    out["residual"] = m.residual(0)
    out["restrict"] = m.restrict(out["residual"])
    out["interpolate"] = m.interpolate(g["in_v"].copy(), g["in_coarse"])
    out["apply_correction"] = m.apply_correction(g["in_v"].copy(), g["in_f"])
    m = make(n, rng_range)
    out["problem_f"] = m.f(0).copy()
    out["problem_v"] = m.v(0).copy()
    hist = [m.residual_norms(0)]
    for _ in range(cycles):
        m.vcycle(0, nu, nu)
        hist.append(m.residual_norms(0))
    out["vcycle_hist"] = np.array(hist)
    out["vcycle_v"] = m.v(0).copy()
    m = make(n, rng_range)
    m.fmg(0, *fmg)
    out["fmg_v"] = m.v(0).copy()
    return out


def run_engine(mg, dim, dtype, corrected, g):
    """The same sequence through the product's C ABI (GPU)."""
    n, rng_range = int(g["n"]), tuple(g["range"])
    nu, cycles, fmg = int(g["nu"]), int(g["cycles"]), tuple(int(x) for x in g["fmg"])
    mode = mg.MG_CORRECTED if corrected else mg.MG_REF_COMPAT

    def make():
        if dim == 3:
            return mg.MultiGrid3D(n, rng_range, dtype=dtype, residual_mode=mode)
        if dim == 2:
            return mg.MultiGrid2D(n, rng_range, dtype=dtype)
        return mg.MultiGrid1D(n, rng_range, dtype=dtype, residual_mode=mode)

    out = {}
    e = make()
    e.set_v(0, g["in_v"])
    e.set_f(0, g["in_f"])
    e.Relax(0, 2)
    out["relax2_v"] = e.get_v(0)
    out["residual"] = e.CalculateResidual(0)
    out["restrict"] = e.Restrict(out["residual"])
    out["interpolate"] = e.Interpolate(g["in_v"].copy(), g["in_coarse"])
    out["apply_correction"] = e.ApplyCorrection(g["in_v"].copy(), g["in_f"])
    e.close()
    e = make()
    out["problem_f"] = e.get_f(0)
    out["problem_v"] = e.get_v(0)
    hist = [e.residual_norm(0)]
    for _ in range(cycles):
        e.VCycle(0, nu, nu)
        hist.append(e.residual_norm(0))
    out["vcycle_hist"] = np.array(hist)
    out["vcycle_v"] = e.get_v(0)
    e.close()
    e = make()
    e.FullMultiGridVCycle(0, *fmg)
    out["fmg_v"] = e.get_v(0)
    e.close()
    return out


def compare(got, g, hist_rtol):
    from util import assert_bits_equal
    for key, val in got.items():
        want = g[key]
        if key == "vcycle_hist":
            assert val.shape == want.shape
            assert np.all(np.abs(val - want) <= hist_rtol * np.abs(want)), (key, val, want)
        else:
            assert_bits_equal(val, want, key)


_GOLD = np.uint64(0x9E3779B97F4A7C15)


def field_checksum(a):
    """Position-keyed additive checksum of a dense field (any dimension, C order = the reference's x-fastest layout):

        sum over points i of mix64(bits(a[i]) + (i + 1) * 0x9E3779B97F4A7C15)   (mod 2^64)

    with bits() the raw IEEE bit pattern zero-extended to 64 bits, i the linear index of the reference layout
    (x + y*n + z*n*n) and mix64 the splitmix64 finaliser.  Addition commutes, so slabs (and GPUs) can be summed in
    any order; the key makes it sensitive to where a value sits.  The engine computes the same number on the device
    (mg3d_field_checksum, csrc/mg3d_kernels.cu)."""
    a = np.ascontiguousarray(a)
    flat = a.reshape(-1)
    bits_t = np.uint32 if a.dtype == np.float32 else np.uint64
    total = np.uint64(0)
    step = 1 << 24
    with np.errstate(over="ignore"):
        for s0 in range(0, flat.size, step):
            b = flat[s0:s0 + step].view(bits_t).astype(np.uint64)
            idx = np.arange(s0 + 1, s0 + 1 + b.size, dtype=np.uint64)
            z = b + idx * _GOLD
            z ^= z >> np.uint64(30)
            z *= np.uint64(0xBF58476D1CE4E5B9)
            z ^= z >> np.uint64(27)
            z *= np.uint64(0x94D049BB133111EB)
            z ^= z >> np.uint64(31)
            total += z.sum(dtype=np.uint64)
    return int(total)
