"""V(2,2) time of the small 3D hierarchies (graph-replayed): the fixed, latency-bound cost every GPU pays for the
agglomerated coarse levels of a slab-partitioned run."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import pde_multigrid_b200 as mg

for n in (9, 17, 33, 65, 129, 257):
    e = mg.MultiGrid3D(n, dtype=np.float64, residual_mode=mg.MG_CORRECTED)
    s = torch.cuda.ExternalStream(e.stream)
    for _ in range(5):
        e.VCycle(0, 2, 2)
    e.sync()
    l0 = e.kernel_launches
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(50):
        e.VCycle(0, 2, 2)
    b.record(s)
    e.sync()
    print("n=%4d  V(2,2) %.1f us  launches/cycle %d" % (n, a.elapsed_time(b) / 50 * 1e3, (e.kernel_launches - l0) // 50))
    e.close()
