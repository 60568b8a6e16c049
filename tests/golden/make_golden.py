#!/usr/bin/env python3
"""tests/golden/make_golden.py -- generates the golden vectors under tests/golden/ by RUNNING THE
REFERENCE ITSELF (oracle/_ref/libmg_ref.so = NOCUDA_TESI compiled unmodified, see oracle/build_ref.py).
Run in a container that has /root/reference:

    python oracle/build_ref.py && python tests/golden/make_golden.py

Each .npz holds seeded inputs and the reference's outputs for one (dimension, dtype, residual mode):
per-operator results on random fields, the residual-norm history of V-cycles on the reference problem
and the final solution, and an FMG solve.  The reference ships no tests or vectors of its own
(SURVEY.md section 4), so these files are the pin for the plain-C oracle (tests/test_oracle.py) and, on
the GPU box, for the CUDA path (tests/test_golden_gpu.py).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref import RefMG  # noqa: E402

CASES = {
    # dim: (n, range, vcycle nu, cycles, fmg (v0, v1, v2))
    3: (17, (0, 1, 0, 1, 0, 1), 2, 3, (2, 3, 3)),
    2: (33, (0, 20, 0, 20), 2, 3, (2, 50, 50)),
    1: (129, (0, 1), 100, 2, (2, 100, 100)),
}


def make(dim, dtype, corrected):
    n, rng_range, nu, cycles, fmg = CASES[dim]
    out = {"n": n, "range": np.array(rng_range, dtype=np.float64), "nu": nu, "cycles": cycles, "fmg": np.array(fmg)}
    rng = np.random.default_rng(12345 + dim)
    v0 = rng.uniform(-1, 1, (n,) * dim).astype(dtype)
    f0 = rng.uniform(-1, 1, (n,) * dim).astype(dtype)
    cn = (n - 1) // 2 + 1
    c0 = rng.uniform(-1, 1, (cn,) * dim).astype(dtype)
    out.update(in_v=v0, in_f=f0, in_coarse=c0)
    # operators on random fields
    m = RefMG(dim, dtype, corrected, n=n, range=rng_range)
    m.v(0)[...] = v0
    m.f(0)[...] = f0
    m.relax(0, 2)
    out["relax2_v"] = m.v(0).copy()
    out["residual"] = m.residual(0)
    out["restrict"] = m.restrict(out["residual"])
    fine = v0.copy()
    m.interpolate(fine, c0)
    out["interpolate"] = fine
    out["apply_correction"] = m.apply_correction(v0.copy(), f0)
    m.close()
    # V-cycle history on the reference problem from v = 0
    m = RefMG(dim, dtype, corrected, n=n, range=rng_range)
    out["problem_f"] = m.f(0).copy()
    out["problem_v"] = m.v(0).copy()
    hist = [m.residual_norms(0)]
    for _ in range(cycles):
        m.vcycle(0, nu, nu)
        hist.append(m.residual_norms(0))
    out["vcycle_hist"] = np.array(hist, dtype=np.float64)
    out["vcycle_v"] = m.v(0).copy()
    m.close()
    m = RefMG(dim, dtype, corrected, n=n, range=rng_range)
    m.fmg(0, *fmg)
    out["fmg_v"] = m.v(0).copy()
    m.close()
    return out


def main():
    for dim in (1, 2, 3):
        for dtype in (np.float32, np.float64):
            for corrected in ((False, True) if dim != 2 else (False,)):
                name = "ref%dd_%s%s.npz" % (dim, "f32" if dtype == np.float32 else "f64", "c" if corrected else "")
                np.savez_compressed(os.path.join(HERE, name), **make(dim, dtype, corrected))
                print(name, os.path.getsize(os.path.join(HERE, name)))


if __name__ == "__main__":
    main()
