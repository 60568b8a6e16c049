#!/usr/bin/env python3
"""Times one level-0 operator with CUDA events on the engine's stream: scripts/bench_op.py rr|interp [n] [dtype] [reps]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import pde_multigrid_b200 as mg  # noqa: E402

op = sys.argv[1] if len(sys.argv) > 1 else "rr"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1025
dt = np.float64 if (len(sys.argv) <= 3 or sys.argv[3] == "f64") else np.float32
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
eng = mg.MultiGrid3D(n, dtype=dt, residual_mode=mg.MG_CORRECTED)
stream = torch.cuda.ExternalStream(eng.stream)
eng.VCycle(0, 2, 2)
fn = {"rr": lambda: eng.residual_restrict(0), "interp": lambda: eng.interpolate_correct(0)}[op]
fn()
eng.sync()
times = []
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    fn()
    b.record(stream)
    eng.sync()
    times.append(a.elapsed_time(b))
print(json.dumps({"op": op, "n": n, "ms_min": min(times), "ms_all": [round(t, 3) for t in times], "env": {k: v for k, v in os.environ.items() if k.startswith("MG_B200")}}))
eng.close()
