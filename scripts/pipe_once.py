#!/usr/bin/env python3
"""Level-0 Relax(0, 2) launches of the chosen smoother at n^3 on a field that a V-cycle has filled (for ncu captures):
python scripts/pipe_once.py [n] [dtype] [smoother] [arith].  The V-cycle runs with MG_SMOOTHER_TMA so that the only
k_relax_pipe2 launches of the process are the level-0 ones."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pde_multigrid_b200 as mg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1025
dt = np.float64 if (len(sys.argv) <= 2 or sys.argv[2] == "f64") else np.float32
sm = sys.argv[3] if len(sys.argv) > 3 else "pipe"
eng = mg.MultiGrid3D(n, dtype=dt, residual_mode=mg.MG_CORRECTED)
eng.set_smoother(mg.MG_SMOOTHER_TMA)
eng.VCycle(0, 2, 2)
eng.set_smoother({"pipe": mg.MG_SMOOTHER_PIPE, "tma": mg.MG_SMOOTHER_TMA}[sm])
if len(sys.argv) > 4 and sys.argv[4] == "fast":
    eng.set_arith(mg.MG_ARITH_FAST)
eng.Relax(0, 2)
eng.Relax(0, 2)
eng.sync()
print("done, residual", eng.residual_norm(0))
eng.close()
