"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py
3D at 257^3 (TMA smoother, fused residual+restrict, cell prolongation, coarse tail, fused 2-sweep smoother),
2D at 257^2 (colour kernels + single-CTA small-level smoother), 1D at 1025 (persistent cycle kernel)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import pde_multigrid_b200 as mg

for dtype in (np.float64, np.float32):
    e = mg.MultiGrid3D(257, dtype=dtype, residual_mode=mg.MG_CORRECTED)
    r0 = e.residual_norm(0)[0]
    e.VCycle(0, 2, 2)
    e.set_smoother(mg.MG_SMOOTHER_FUSED, 2)
    e.VCycle(0, 2, 2)
    e.set_smoother(mg.MG_SMOOTHER_COLOUR, 1)
    e.VCycle(0, 1, 1)
    e.FullMultiGridVCycle(0, 1, 1, 1)
    r1 = e.residual_norm(0)[0]
    v = e.get_v(0)
    _ = e.CalculateResidual(0)
    c = e.Restrict(v)
    e.Interpolate(v, c)
    assert np.isfinite(r1) and r1 < r0
    e.close()
    e2 = mg.MultiGrid2D(257, dtype=dtype)
    e2.VCycle(0, 2, 2)
    e2.FullMultiGridVCycle(0, 1, 20, 20)
    assert np.isfinite(e2.residual_norm(0)[0]) and np.isfinite(e2.mean_abs_error())
    e2.close()
    e1 = mg.MultiGrid1D(1025, dtype=dtype)
    e1.VCycle(0, 50, 50)
    e1.FullMultiGridVCycle(0, 1, 50, 50)
    assert np.isfinite(e1.residual_norm(0)[0])
    e1.close()
print("sanitize_smoke ok")
