"""Time of the residual-norm pass at 1025^3 fp64 (CUDA events on the engine's stream)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import pde_multigrid_b200 as mg

e = mg.MultiGrid3D(1025, dtype=np.float64, residual_mode=mg.MG_CORRECTED)
e.VCycle(0, 2, 2)
s = torch.cuda.ExternalStream(e.stream)
print("norm", e.residual_norm(0))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(s)
for _ in range(10):
    r = e.residual_norm(0)
b.record(s)
e.sync()
print("residual_norm: %.3f ms per call (17.2 GB algorithmic: %.0f GB/s)" % (a.elapsed_time(b) / 10, 17.23 / (a.elapsed_time(b) / 10) * 1e3))
