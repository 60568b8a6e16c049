#!/usr/bin/env python3
"""Times Relax(level 0, nu) for each smoother implementation with CUDA events on the engine's stream.

    python scripts/bench_smoother.py [--n 1025] [--dtype f64] [--nu 2] [--reps 5] [--smoothers tma,pipe,pipe_fast]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1025)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--nu", type=int, default=2)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--smoothers", default="tma,pipe,pipe_fast")
    args = ap.parse_args()
    import torch
    import pde_multigrid_b200 as mg
    dt = np.float64 if args.dtype == "f64" else np.float32
    B = np.dtype(dt).itemsize
    eng = mg.MultiGrid3D(args.n, dtype=dt, residual_mode=mg.MG_CORRECTED)
    stream = torch.cuda.ExternalStream(eng.stream)
    eng.VCycle(0, 2, 2)  # something other than zeros in v
    for name in args.smoothers.split(","):
        code = {"tma": mg.MG_SMOOTHER_TMA, "pipe": mg.MG_SMOOTHER_PIPE, "pipe_fast": mg.MG_SMOOTHER_PIPE,
                "colour": mg.MG_SMOOTHER_COLOUR, "fused": mg.MG_SMOOTHER_FUSED}[name]
        eng.set_smoother(code)
        eng.set_arith(mg.MG_ARITH_FAST if name == "pipe_fast" else mg.MG_ARITH_EXACT)
        eng.Relax(0, args.nu)
        eng.sync()
        times = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            eng.Relax(0, args.nu)
            e1.record(stream)
            eng.sync()
            times.append(e0.elapsed_time(e1))
        ms = min(times)
        N = args.n ** 3
        print(json.dumps({"smoother": name, "n": args.n, "dtype": args.dtype, "nu": args.nu, "ms_min": ms,
                          "ms_all": [round(t, 3) for t in times],
                          "algorithmic_GBps": args.nu * 3 * B * N / (ms * 1e-3) / 1e9}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
