"""Quick device timing of the 3D V-cycle (development aid, not the bench)."""
import sys, time, ctypes
import numpy as np
sys.path.insert(0, ".")
import torch
import pde_multigrid_b200 as mg

def run(n, dtype, cycles=5, nu=2):
    eng = mg.MultiGrid3D(n, dtype=dtype, residual_mode=mg.MG_CORRECTED)
    s = torch.cuda.ExternalStream(eng.stream)
    eng.VCycle(0, nu, nu); eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.kernel_launches
    with torch.cuda.stream(s):
        e0.record(s)
        for _ in range(cycles):
            eng.VCycle(0, nu, nu)
        e1.record(s)
    eng.sync()
    ms = e0.elapsed_time(e1) / cycles
    B = np.dtype(dtype).itemsize
    sizes = [eng.level_size(l) for l in range(eng.numGrids)]
    N = [s_ ** 3 for s_ in sizes]
    byt = sum(2 * nu * 3 * B * x for x in N) + sum(4 * B * N[l] + 3 * B * N[l + 1] for l in range(len(N) - 1))
    print("n=%d %s: %.3f ms/cycle  %.1f cyc/s  %.0f GB/s algorithmic (%.1f%% of 6451.8)  launches/cycle %d  norm %s" % (
        n, np.dtype(dtype).name, ms, 1e3 / ms, byt / ms / 1e6, 100 * byt / ms / 1e6 / 6451.8,
        (eng.kernel_launches - l0) // cycles, eng.residual_norm(0)))
    eng.close()

for n in [int(a) for a in sys.argv[1:]] or [129, 257, 513]:
    run(n, np.float64)
    run(n, np.float32)
