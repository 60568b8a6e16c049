#!/usr/bin/env python3
"""One eager V(2,2) cycle at n^3 with the default smoother (for ncu captures of kernels in their V-cycle context):
python scripts/vcycle_once.py [n] [dtype]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pde_multigrid_b200 as mg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1025
dt = np.float64 if (len(sys.argv) <= 2 or sys.argv[2] == "f64") else np.float32
eng = mg.MultiGrid3D(n, dtype=dt, residual_mode=mg.MG_CORRECTED)
eng.VCycle(0, 2, 2)
eng.sync()
print("done", eng.residual_norm(0))
eng.close()
