"""Small end-to-end exercise of every kernel family, for a checking build of the library:
    MG_B200_LIB=pde_multigrid_b200/libmg_b200_dbg.so python scripts/sanitize_smoke.py      (-DMG_DEBUG_BOUNDS, tests/test_debug_bounds.py)
    compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py                     (where the tool is available)
3D at 257^3 and 513^3 (temporally blocked smoother with and without the fused prolongation, its literal-arithmetic fallback,
TMA colour smoother, fused residual+restrict, cell prolongation, coarse tail, Jacobi, FMG, host-array operators, diagnostics),
2D at 257^2 (colour kernels + single-CTA small-level smoother), 1D at 1025 (persistent cycle kernel)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import pde_multigrid_b200 as mg

for dtype in (np.float64, np.float32):
    for n in (257, 513) if dtype == np.float64 else (257,):
        e = mg.MultiGrid3D(n, dtype=dtype, residual_mode=mg.MG_CORRECTED)
        r0 = e.residual_norm(0)[0]
        e.VCycle(0, 2, 2)   # default: two-sweep passes, prolongation folded into the post-smoothing pass
        e.VCycle(0, 2, 2)   # captured
        e.VCycle(0, 2, 2)   # replayed
        e.VCycle(0, 3, 1)   # pass + colour remainder, separate prolongation
        e.set_smoother(mg.MG_SMOOTHER_FUSED)
        e.VCycle(0, 2, 2)
        e.set_smoother(mg.MG_SMOOTHER_TMA)
        e.VCycle(0, 1, 1)
        e.set_smoother(mg.MG_SMOOTHER_COLOUR)
        e.VCycle(0, 1, 1)
        e.set_smoother(mg.MG_SMOOTHER_JACOBI)
        e.VCycle(0, 2, 1)
        e.set_smoother(mg.MG_SMOOTHER_AUTO)
        e.FullMultiGridVCycle(0, 1, 2, 2)
        r1 = e.residual_norm(0)[0]
        assert np.isfinite(r1) and r1 < r0
        e.field_checksum(0)
        e.abs_error(0)
        if n == 257:
            v = e.get_v(0)
            v[n // 2, n // 2, 1:9] = np.finfo(dtype).tiny  # the range guard fires: conditional literal-arithmetic pass
            e.set_v(0, v)
            e.VCycle(0, 2, 2)
            _ = e.CalculateResidual(0)
            c = e.Restrict(v)
            e.Interpolate(v, c)
            e.ApplyCorrection(v, v.copy())
        e.close()
    e2 = mg.MultiGrid2D(257, dtype=dtype)
    e2.VCycle(0, 2, 2)
    e2.FullMultiGridVCycle(0, 1, 20, 20)
    assert np.isfinite(e2.residual_norm(0)[0]) and np.isfinite(e2.mean_abs_error())
    e2.close()
    e1 = mg.MultiGrid1D(1025, dtype=dtype)
    e1.VCycle(0, 50, 50)
    e1.FullMultiGridVCycle(0, 1, 50, 50)
    assert np.isfinite(e1.residual_norm(0)[0]) and np.isfinite(e1.abs_error(0)[1])
    e1.close()
print("sanitize_smoke ok")
