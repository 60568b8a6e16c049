"""Throughput of the other BASELINE.json configurations (they are parity-test cases for bench.py, measured here
for the record): configs[0] 1D N=1025, configs[1] 2D Lyapunov 1025^2, configs[2] 3D Poisson 257^3 fp64.
Prints one JSON line per measurement.  CUDA events on the engine's stream; the CPU column is the reference's
own NOCUDA_TESI solver (oracle/_ref, 1 thread) on this box's host when --cpu is given."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import pde_multigrid_b200 as mg


def timed(eng, fn, reps, warm=2):
    s = torch.cuda.ExternalStream(eng.stream)
    for _ in range(warm):
        fn()
    eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(reps):
        fn()
    e1.record(s)
    eng.sync()
    return e0.elapsed_time(e1) / reps


def cpu_time(make, fn, reps=1):
    o = make()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn(o)
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    with_cpu = "--cpu" in sys.argv
    if with_cpu:
        from oracle import ref, port
        Or = ref.RefMG if ref.available() else port.PortMG
    out = []
    # configs[0]: 1D, N = 1025, thesis parameters nu = 1000 (N1/Poisson1DSolver.cpp:15-25)
    for dtype in (np.float32, np.float64):
        e = mg.MultiGrid1D(1025, dtype=dtype)
        ms_v = timed(e, lambda: e.VCycle(0, 1000, 1000), 5)
        ms_f = timed(e, lambda: (e.init_problem(), e.FullMultiGridVCycle(0, 2, 1000, 1000)), 3)
        rec = {"config": "1D N=1025 %s" % np.dtype(dtype).name, "vcycle_1000_1000_ms": ms_v, "fmg_2_1000_1000_ms": ms_f,
               "launches_per_vcycle": 1}
        if with_cpu:
            rec["cpu_vcycle_ms"] = cpu_time(lambda: Or(1, dtype, False, n=1025), lambda o: o.vcycle(0, 1000, 1000))
            rec["cpu_fmg_ms"] = cpu_time(lambda: Or(1, dtype, False, n=1025), lambda o: o.fmg(0, 2, 1000, 1000))
        out.append(rec)
        e.close()
    # configs[1]: 2D Lyapunov 1025^2 (fp32 is the reference's native type)
    for dtype in (np.float32, np.float64):
        e = mg.MultiGrid2D(1025, dtype=dtype)
        l0 = e.kernel_launches
        ms_v = timed(e, lambda: e.VCycle(0, 2, 2), 20)
        lv = (e.kernel_launches - l0) // 22
        ms_f = timed(e, lambda: (e.init_problem(), e.FullMultiGridVCycle(0, 1, 500, 500)), 2, warm=1)
        upd = 4 * sum((s - 2) ** 2 for s in [1025, 513, 257, 129, 65, 33, 17, 9, 5, 3])
        rec = {"config": "2D Lyapunov 1025^2 %s" % np.dtype(dtype).name, "vcycle_2_2_ms": ms_v, "vcycles_per_s": 1e3 / ms_v,
               "grid_point_updates_per_s": upd * 1e3 / ms_v, "fmg_1_500_500_ms": ms_f, "launches_per_vcycle": lv}
        if with_cpu:
            rec["cpu_vcycle_ms"] = cpu_time(lambda: Or(2, dtype, False, n=1025), lambda o: o.vcycle(0, 2, 2), 3)
        out.append(rec)
        e.close()
    # configs[2]: 3D Poisson 257^3 fp64 V(2,2)
    for dtype in (np.float64, np.float32):
        e = mg.MultiGrid3D(257, dtype=dtype, residual_mode=mg.MG_CORRECTED)
        ms_v = timed(e, lambda: e.VCycle(0, 2, 2), 50, warm=3)
        B = np.dtype(dtype).itemsize
        sizes = [257, 129, 65, 33, 17, 9, 5, 3]
        N = [s ** 3 for s in sizes]
        byt = sum(12 * B * x for x in N) + sum(4 * B * N[l] + 3 * B * N[l + 1] for l in range(len(N) - 1))
        upd = 4 * sum((s - 2) ** 3 for s in sizes)
        rec = {"config": "3D Poisson 257^3 %s V(2,2)" % np.dtype(dtype).name, "vcycle_ms": ms_v, "vcycles_per_s": 1e3 / ms_v,
               "grid_point_updates_per_s": upd * 1e3 / ms_v, "algorithmic_gbs": byt / ms_v / 1e6}
        if with_cpu and dtype == np.float64:
            rec["cpu_vcycle_ms"] = cpu_time(lambda: Or(3, dtype, True, n=257), lambda o: o.vcycle(0, 2, 2))
        out.append(rec)
        e.close()
    # thesis replay: FullMultiGridVCycle(0, v0=2, nu=3000, nu=3000) at 257^3 in float, the run the
    # thesis reports at 294 s on its GPU (T:p.69-72).  The thesis code carries the residual sign defect (REF_COMPAT);
    # the memory traffic of the two modes is identical.
    for mode, name in ((mg.MG_REF_COMPAT, "ref_compat"), (mg.MG_CORRECTED, "corrected")):
        e = mg.MultiGrid3D(257, dtype=np.float32, residual_mode=mode)
        s = torch.cuda.ExternalStream(e.stream)
        e.init_problem()
        e.sync()
        l0 = e.kernel_launches
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        e.FullMultiGridVCycle(0, 2, 3000, 3000)
        b.record(s)
        e.sync()
        out.append({"config": "thesis replay: 3D Poisson 257^3 float32 FMG(v0=2, nu1=nu2=3000), %s" % name,
                    "seconds": a.elapsed_time(b) / 1e3, "launches": int(e.kernel_launches - l0), "thesis_gpu_seconds": 294.0})
        e.close()
    # the reference's entry point at the headline size: FullMultiGridVCycle(0, 1, 2, 2) on 1025^3 fp64
    e = mg.MultiGrid3D(1025, dtype=np.float64, residual_mode=mg.MG_CORRECTED)
    r0 = e.residual_norm(0)[0]
    ms_f = timed(e, lambda: (e.init_problem(), e.FullMultiGridVCycle(0, 1, 2, 2)), 3, warm=1)
    r1 = e.residual_norm(0)[0]
    out.append({"config": "3D Poisson 1025^3 float64 init + FMG(v0=1, 2, 2)", "ms": ms_f, "residual_l2_before": r0, "residual_l2_after": r1})
    e.close()
    # weighted-Jacobi option, 257^3 fp64 V(2,2)
    e = mg.MultiGrid3D(257, dtype=np.float64, residual_mode=mg.MG_CORRECTED)
    e.set_smoother(mg.MG_SMOOTHER_JACOBI)
    out.append({"config": "3D Poisson 257^3 float64 V(2,2), weighted Jacobi (omega = 6/7)", "vcycle_ms": timed(e, lambda: e.VCycle(0, 2, 2), 50, warm=3)})
    e.close()
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
